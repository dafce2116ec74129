#!/bin/bash
# go 9x9 MCTS: effect of the history filter and of the float32 pre-selection (B2S_MCTS_TUNING bit 0 / bit 1 switch them off)
mkdir -p gpurun_out
OUT=gpurun_out/${1:-r02}_mcts_variants.jsonl
: > $OUT
for t in 0 1 2 3; do
  for cfg in "16384 256" "8192 4000" "4096 10000"; do
    set -- $cfg
    B2S_MCTS_TUNING=$t python scripts/bench_mcts.py $1 $2 | tail -1 | sed "s/^{/{\"tuning\": $t, /" | tee -a $OUT
  done
done
