#!/usr/bin/env python3
"""Launch one streaming kernel a few times at a given batch size (for ncu): profile_one.py <kernel> <log2 n> [game]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
import open_spiel_b200 as b2  # noqa: E402

kernel, logn = sys.argv[1], int(sys.argv[2])
gs = sys.argv[3] if len(sys.argv) > 3 else "connect_four"
dev = torch.device("cuda", 0)
n = 1 << logn
game = b2.load_game(gs)
if gs == "connect_four":
    _, snap, actions = bench.build_workload(torch, game, n, dev, seed=1)
else:
    import sweep_games                                   # mid-game states + one legal action per lane, as the sweep builds them
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    snap = game.new_batch(n)
    for _ in range(dict(sweep_games.GAMES)[gs]):
        snap.apply_actions(sweep_games.random_legal(snap.legal_actions_mask_words(), gen))
    actions = sweep_games.random_legal(snap.legal_actions_mask_words(), gen)
    snap.check_errors()
work = game.new_batch(n)
mask = torch.empty((n, game._info.mask_words), dtype=torch.int32, device=dev)
term = torch.empty((n,), dtype=torch.uint8, device=dev)
rets = torch.empty((n, 2), dtype=torch.float32, device=dev)
obs = torch.empty((min(n, 1 << 20), game.observation_tensor_size()), dtype=torch.float32, device=dev)
fns = {"apply": lambda: work.apply_actions(actions), "step_fused": lambda: work.step(actions, mask, term, rets),
       "legal_mask": lambda: work.legal_actions_mask_words(out=mask), "status": lambda: work.status(),
       "observation": lambda: work.observation_tensor(0, out=obs, n=obs.shape[0])}
for i in range(4):
    work.copy_from(snap)
    fns[kernel]()
torch.cuda.synchronize()
print("done", kernel, n)
