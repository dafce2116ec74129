set -x
cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_parity_games.py tests/test_gpu_cpp_adapter.py -x -q 2>&1 | tail -8
timeout 600 python scripts/e2e_chunks.py 2>&1 | tail -8
