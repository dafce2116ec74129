# Multi-GPU round-2 checks (gpurun --gpus N -- bash scripts/r02_gpu_dist.sh N): host-side protocol check, the exact sharded CFR
# measurement, and the bench at N GPUs (both arms).  Output under gpurun_out/r02_*_${N}gpu*.
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02_gpus_${N}.csv
nvidia-smi topo -m > gpurun_out/r02_topo_${N}.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist_check"
timeout 600 $TR --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -8
echo "== cfr dist"
timeout 900 $TR --master-port 29531 scripts/r02_cfr_dist.py gpurun_out/r02_cfr_dist_${N}gpu.json 2> gpurun_out/r02_cfr_dist_${N}.err | tail -1 | cut -c1-1500
tail -3 gpurun_out/r02_cfr_dist_${N}.err
echo "== bench reference arm"
timeout 900 $TR --master-port 29513 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02_bench_ref_${N}.err | tail -1 > gpurun_out/r02_bench_ref_${N}gpu.json
cut -c1-400 gpurun_out/r02_bench_ref_${N}gpu.json
echo "== bench"
timeout 1200 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02_bench_${N}.err | tail -1 > gpurun_out/r02_bench_${N}gpu.json
tail -3 gpurun_out/r02_bench_${N}.err
python - <<P
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["roofline"]["frac"])
print(json.dumps(d["config"].get("loops_summary")))
P
if [ "$N" = "2" ]; then
echo "== single-GPU tests touched since the last full run"
timeout 900 python -m pytest tests/test_gpu_cpp_adapter.py tests/test_gpu_parity_games.py tests/test_gpu_vs_reference.py tests/test_gpu_pyspiel.py -x -q -m gpu 2>&1 | tail -5
else
timeout 300 python -m pytest tests/test_gpu_parity_games.py tests/test_gpu_vs_reference.py -q -m gpu -k "leduc or mnk" 2>&1 | tail -3
fi
