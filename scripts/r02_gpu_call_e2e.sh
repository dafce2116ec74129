cd /root/repo
mkdir -p gpurun_out
echo "== host-step tests"
timeout 900 python -m pytest tests/test_gpu_parity_games.py tests/test_gpu_bench_workload.py tests/test_gpu_trajectories.py -x -q -m gpu -k "host or bench or traj" -s 2>&1 | grep -v "^$" | tail -6
echo "== e2e by configuration"
timeout 600 python scripts/r02_e2e_graph.py 2>&1 | tee gpurun_out/r02_e2e_graph.txt
echo "== bench K=20"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_bench_k20b.err | tail -1 > gpurun_out/r02_bench_k20b.json
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_k20b.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"])
P
tail -2 gpurun_out/r02_bench_k20b.err
