set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_mccfr.py tests/test_gpu_serialization.py tests/test_gpu_parity_games.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'P'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"])
print(json.dumps(d["extras"]["loops"], indent=0))
P
tail -3 gpurun_out/bench_quick.err
