#!/usr/bin/env python3
"""Per-game timing of the streaming kernels (device-resident, L2 flushed by a 256 MiB write between launches):
sweep_games.py [log2 lanes]   -> one JSON line per (game, kernel)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import open_spiel_b200 as b2  # noqa: E402

GAMES = [("tic_tac_toe", 3), ("connect_four", 10), ("breakthrough", 12), ("hex", 30), ("go(board_size=9)", 40),
         ("kuhn_poker", 2), ("leduc_poker", 3)]


def random_legal(mask_words, gen):
    n, W = mask_words.shape
    bits = torch.arange(32, device=mask_words.device, dtype=torch.int32)
    dense = ((mask_words.unsqueeze(-1) >> bits) & 1).reshape(n, W * 32).bool()
    score = torch.rand((n, W * 32), device=mask_words.device, generator=gen).masked_fill(~dense, -1.0)
    a = score.argmax(dim=1).to(torch.int32)
    return torch.where(dense.any(dim=1), a, torch.full_like(a, -1))


def main():
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << logn
    dev = torch.device("cuda", 0)
    peak, _ = bench.hbm_peak()
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for gs, plies in GAMES:
        game = b2.load_game(gs)
        m = n if game._info.mask_words <= 4 else min(n, 1 << 18)       # dense sampling needs n*W*32 floats
        snap = game.new_batch(m)
        for _ in range(plies):
            snap.apply_actions(random_legal(snap.legal_actions_mask_words(), gen))
        actions = random_legal(snap.legal_actions_mask_words(), gen)
        snap.check_errors()
        work = game.new_batch(m)
        info = game._info
        sb = info.state_bytes
        mask = torch.empty((m, info.mask_words), dtype=torch.int32, device=dev)
        term = torch.empty((m,), dtype=torch.uint8, device=dev)
        rets = torch.empty((m, 2), dtype=torch.float32, device=dev)
        obs = torch.empty((min(m, 1 << 18), info.observation_tensor_size), dtype=torch.float32, device=dev)
        kernels = {
            "apply": (lambda: work.apply_actions(actions), 2 * sb + 4, m),
            "step_fused": (lambda: work.step(actions, mask, term, rets), 2 * sb + 4 + 9 + 4 * info.mask_words, m),
            "legal_mask": (lambda: work.legal_actions_mask_words(out=mask), sb + 4 * info.mask_words, m),
            "status": (lambda: work.status(), sb + 10, m),
            "observation": (lambda: work.observation_tensor(0, out=obs, n=obs.shape[0]), sb + 4 * info.observation_tensor_size, obs.shape[0]),
        }
        for name, (fn, bytes_per, units) in kernels.items():
            ts = []
            for i in range(9):
                work.copy_from(snap)
                flush.fill_(i)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            gbs = bytes_per * units / (ms / 1e3) / 1e9
            print(json.dumps({"game": gs, "kernel": name, "lanes": units, "state_bytes": sb, "bytes_per_lane": bytes_per,
                              "ms": round(ms, 5), "per_s": units / (ms / 1e3), "alg_GBps": round(gbs, 1),
                              "frac_of_peak": round(gbs / peak, 3)}), flush=True)
        del work, snap


if __name__ == "__main__":
    main()
