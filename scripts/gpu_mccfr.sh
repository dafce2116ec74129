set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_mccfr.py -x -q 2>&1 | tail -15
timeout 600 python scripts/bench_mccfr.py 2>&1 | tee gpurun_out/mccfr_bench.txt | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_mccfr -c 30 --csv --log-file gpurun_out/mccfr_launches.csv python - <<'P'
import sys
sys.path.insert(0, ".")
import open_spiel_b200 as b2
s = b2.ExternalSamplingMCCFRSolver(b2.load_game("leduc_poker"), seed=1, traversals_per_update=65536)
s.run_iteration(6)
P
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
