#!/bin/bash
# ncu source-level profile of the go 9x9 MCTS kernel, shallow (4096 trees x 256 sims) and at depth (1024 trees x 4000 sims).
# Limited section set (a few replay passes): per-line instruction counts (-lineinfo), warp stall states, occupancy.
set -u
mkdir -p gpurun_out
SECS="--section SourceCounters --section WarpStateStats --section InstructionStats --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats"
timeout 900 ncu $SECS --import-source on --clock-control none -k regex:k_mcts -c 1 -f -o gpurun_out/${1:-r02}_prof_mcts_shallow \
    python scripts/bench_mcts.py 4096 256 > gpurun_out/${1:-r02}_prof_mcts_shallow.log 2>&1
timeout 1500 ncu $SECS --import-source on --clock-control none -k regex:k_mcts -c 1 -f -o gpurun_out/${1:-r02}_prof_mcts_deep \
    python scripts/bench_mcts.py 1024 4000 > gpurun_out/${1:-r02}_prof_mcts_deep.log 2>&1
ls -la gpurun_out | grep prof_mcts
