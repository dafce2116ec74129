#!/bin/bash
mkdir -p gpurun_out
echo "== cuda-gdb adapter_test_dbg"
timeout 600 /usr/local/cuda/bin/cuda-gdb -batch -ex run -ex bt -ex "info threads" --args open_spiel_b200/adapter/_build/adapter_test_dbg > gpurun_out/r02_adapter_gdb.log 2>&1
tail -40 gpurun_out/r02_adapter_gdb.log
echo "== pytest (cfr, adapter, mcts, mccfr, pyspiel)"
timeout 1200 python -m pytest tests -q -m gpu -k "cfr or adapter or mcts or mccfr or pyspiel" 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu_partial.log
echo "== mcts quick"
for cfg in "16384 256" "65536 128" "8192 4000" "8192 10000"; do set -- $cfg; python scripts/bench_mcts.py $1 $2 | tail -1 | tee -a gpurun_out/r02_mcts_after_filter.jsonl; done
echo "== mcts 100k sims/move (BASELINE configs[2] depth), 8192 trees"
timeout 900 python scripts/bench_mcts.py 8192 100000 | tail -1 | tee gpurun_out/r02_mcts_go9x9_100k_sims.json
