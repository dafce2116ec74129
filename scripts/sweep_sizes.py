#!/usr/bin/env python3
"""Batch-size sweep of the connect_four streaming kernels (device-resident, L2 flushed between launches).
Prints one JSON line per (kernel, n): time per launch (CUDA events), steps/s, algorithmic GB/s, fraction of peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import open_spiel_b200 as b2  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    peak, _ = bench.hbm_peak()
    game = b2.load_game("connect_four")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for logn in (16, 18, 20, 22, 24, 26):
        n = 1 << logn
        _, snap, actions = bench.build_workload(torch, game, n, dev, seed=1)
        work = game.new_batch(n)
        mask = torch.empty((n, 1), dtype=torch.int32, device=dev)
        term = torch.empty((n,), dtype=torch.uint8, device=dev)
        rets = torch.empty((n, 2), dtype=torch.float32, device=dev)
        obs = torch.empty((min(n, 1 << 22), 126), dtype=torch.float32, device=dev)
        kernels = {
            "apply": (lambda: work.apply_actions(actions), 36, n),
            "step_fused": (lambda: work.step(actions, mask, term, rets), 49, n),
            "legal_mask": (lambda: work.legal_actions_mask_words(out=mask), 20, n),
            "status": (lambda: work.status(), 26, n),
            "observation": (lambda: work.observation_tensor(0, out=obs, n=obs.shape[0]), 520, obs.shape[0]),
        }
        for name, (fn, bytes_per, units) in kernels.items():
            ts = []
            for i in range(13):
                work.copy_from(snap)
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            gbs = bytes_per * units / (ms / 1e3) / 1e9
            print(json.dumps({"kernel": name, "n": units, "ms": round(ms, 5), "per_s": units / (ms / 1e3),
                              "alg_GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3)}), flush=True)
        del work, snap


if __name__ == "__main__":
    main()
