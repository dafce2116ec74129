"""Throughput of the batched trajectory recorder (b2s_record_trajectories).  Usage: python scripts/bench_traj.py"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import open_spiel_b200 as b2


def run(gs, n, obs, reps=3):
    game = b2.load_game(gs)
    batch = game.new_batch(n)
    best = None
    for r in range(reps + 1):
        batch.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr = batch.record_trajectories(seed=r, include_full_observations=obs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if r > 0 and (best is None or ms < best):
            best = ms
        steps = int(tr.lengths.sum())
        T = tr.max_trajectory_length
        bytes_out = sum(v.numel() * v.element_size() for v in tr.time_major.values() if v is not None)
        del tr
    return {"game": gs, "episodes": n, "observations": obs, "T": T, "ms": round(best, 3),
            "episodes_per_s": n / best * 1e3, "decisions_per_s": steps / best * 1e3,
            "output_GB": bytes_out / 1e9, "output_GBps": bytes_out / best / 1e6}


if __name__ == "__main__":
    assert torch.cuda.is_available()
    for gs, n in [("connect_four", 1 << 20), ("tic_tac_toe", 1 << 22), ("breakthrough", 1 << 16), ("go(board_size=9)", 1 << 16),
                  ("leduc_poker", 1 << 22), ("kuhn_poker", 1 << 22)]:
        for obs in (False, True):
            print(json.dumps(run(gs, n, obs)), flush=True)
