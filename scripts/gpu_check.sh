#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture.
# Usage (from repo root on the GPU box): bash scripts/gpu_check.sh [quick]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02_smoke.log
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/r02_bench_ref.json
echo "== bench (driver's command)"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/r02_bench_k20.json | cut -c1-600
echo "== bench"; timeout 1500 python bench.py --steps 200 --warmup 10 2>&1 | tail -3 | tee gpurun_out/r02_bench.json | cut -c1-2500
echo "== ncu (zero-copy host step)"
cat > /tmp/zc1.py <<'P'
import sys
sys.path.insert(0, ".")
import torch
import open_spiel_b200 as b2
n = 1 << 20
game = b2.load_game("connect_four")
batch = game.new_batch(n)
a8 = torch.randint(0, 7, (n,), dtype=torch.int32).to(torch.uint8).pin_memory()
status = torch.empty((n,), dtype=torch.uint8).pin_memory()
for _ in range(4):
    batch.step_host_compact(a8, status)
torch.cuda.synchronize()
P
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step_compact_zc -s 1 -c 2 -f -o gpurun_out/r02_prof_step_zero_copy python /tmp/zc1.py > /dev/null 2>&1
if [ "${1:-}" != "quick" ]; then
  echo "== ncu launches"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 2000 --csv --log-file gpurun_out/r02_launches.csv \
      python bench.py --steps 20 --warmup 3 --deep-trees 1024 --deep-sims 500 --cfr-iters 2000 > gpurun_out/r02_bench_under_ncu.log 2>&1
  echo "== ncu full (k_apply)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 60 -c 3 -f -o gpurun_out/r02_prof_apply \
      python bench.py --steps 20 --warmup 3 --deep-trees 1024 --deep-sims 500 --cfr-iters 2000 > gpurun_out/r02_bench_under_ncu2.log 2>&1
  ls -la gpurun_out
fi
