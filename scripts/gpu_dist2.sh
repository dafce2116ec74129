set -x
cd /root/repo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 2> gpurun_out/bench2.err | tail -1 > gpurun_out/bench_2gpu_quick.json
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_2gpu_quick.json"))
print(d["n_gpus"], d["value"], d["e2e"]["value"])
print(json.dumps(d["extras"]["loops"]))
P
tail -3 gpurun_out/bench2.err
