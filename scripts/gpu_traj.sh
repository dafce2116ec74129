set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_trajectories.py tests/test_gpu_mcts.py -x -q 2>&1 | tail -15
timeout 300 python scripts/bench_traj.py 2>&1 | tee gpurun_out/traj_bench.txt | tail -20
