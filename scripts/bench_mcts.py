#!/usr/bin/env python3
"""MCTS throughput (BASELINE configs[2]): go(board_size=9) MCTS, RandomRolloutEvaluator(n_rollouts=1), uct_c=2,
solve=true, R independent trees per GPU from the initial position.  sims/s = sum(sims_run) / wall time of
b2s_mcts_search (synchronous).  Usage: bench_mcts.py [trees] [sims] [game]"""
import json
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import open_spiel_b200 as b2  # noqa: E402


def run(game_string, trees, sims, seed=1, nodes=0):
    game = b2.load_game(game_string)
    batch = game.new_batch(trees)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = b2.mcts_search(batch, sims, uct_c=2.0, n_rollouts=1, solve=True, seed=seed, max_nodes_total=nodes)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    total = int(out["sims_run"].sum().item())
    errs = batch.error_count()[0]
    return {"game": game_string, "trees": trees, "sims_per_tree": sims, "seconds": round(dt, 4),
            "sims_per_s": total / dt, "nodes_used": b2.mcts_nodes_used(batch), "errors": errs}


if __name__ == "__main__":
    gs = sys.argv[3] if len(sys.argv) > 3 else "go(board_size=9)"
    if len(sys.argv) > 2:
        print(json.dumps(run(gs, int(sys.argv[1]), int(sys.argv[2]))))
    else:
        run(gs, 256, 16)          # warm-up (module load, allocations)
        for trees, sims in [(1024, 256), (4096, 256), (16384, 256), (65536, 128), (4096, 2048), (16384, 1024)]:
            print(json.dumps(run(gs, trees, sims)), flush=True)
        ref = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
        if os.path.exists(ref):
            for th in (1, 16, 64):
                out = subprocess.run([ref, "mcts", gs, "2000", "1", str(th)], capture_output=True, text=True)
                print("reference CPU MCTSBot:", out.stdout.strip(), flush=True)
