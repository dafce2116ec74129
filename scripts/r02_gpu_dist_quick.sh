# bench at N GPUs only (final code): gpurun --gpus N -- bash scripts/r02_gpu_dist_quick.sh N
N=${1:-2}
cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02_bench_${N}_final.err | tail -1 > gpurun_out/r02_bench_${N}gpu_final.json
tail -3 gpurun_out/r02_bench_${N}_final.err
python - <<P
import json
d = json.load(open("gpurun_out/r02_bench_${N}gpu_final.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"])
print(d["e2e"])
print(json.dumps(d["config"].get("loops_summary")))
P
