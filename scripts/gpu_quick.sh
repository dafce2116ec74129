#!/bin/bash
# quick GPU pass: parity tests + size sweep (+ optional extra command)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 600 python scripts/sweep_sizes.py 2>&1 | tee gpurun_out/sweep.jsonl | tail -40
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 10 2>&1 | tail -3 | tee gpurun_out/bench.json
