set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_mccfr.py -x -q 2>&1 | tail -15
timeout 600 python scripts/bench_mccfr.py 2>&1 | tee gpurun_out/mccfr_bench.txt | tail -20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_mccfr -c 60 --csv --log-file gpurun_out/mccfr_launches.csv python - <<'P'
import sys
sys.path.insert(0, ".")
import open_spiel_b200 as b2
s = b2.ExternalSamplingMCCFRSolver(b2.load_game("leduc_poker"), seed=1, traversals_per_update=65536)
s.run_iteration(6)
s = b2.ExternalSamplingMCCFRSolver(b2.load_game("leduc_poker"), seed=1, traversals_per_update=4096)
s.run_iteration(6)
P
tail -30 gpurun_out/mccfr_launches.csv | cut -c1-200
