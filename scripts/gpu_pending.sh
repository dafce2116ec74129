#!/bin/bash
# experiments waiting for a GPU slot
timeout 600 python -m pytest tests/test_gpu_cfr.py -x -q 2>&1 | tail -4
python - <<PY
import time, torch, sys
sys.path.insert(0, ".")
import open_spiel_b200 as b2
for gs in ("leduc_poker", "kuhn_poker"):
    s = b2.CFRSolver(b2.load_game(gs)); s.evaluate_and_update_policy(10); torch.cuda.synchronize()
    t0 = time.perf_counter(); s.evaluate_and_update_policy(5000); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(gs, "iters/s", 5000 / dt, "exploitability", s.exploitability())
PY
python scripts/bench_mcts.py 16384 256; python scripts/bench_mcts.py 65536 128; python scripts/bench_mcts.py 131072 64; python scripts/bench_mcts.py 262144 48
python scripts/sweep_games.py 22 2>&1 | grep -E "tic_tac|kuhn|leduc" | cut -c1-200
