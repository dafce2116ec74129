"""Host-buffer step (1M connect_four lanes): zero-copy kernel on / off, graph replay on / off, by chunk count, both entry points.
Each configuration in a fresh process (the switches are read once): python scripts/r02_e2e_graph.py"""
import os
import subprocess
import sys

CODE = r'''
import os, sys, time, torch
sys.path.insert(0, ".")
import open_spiel_b200 as b2
from open_spiel_b200 import _lib
n = 1 << 20
game = b2.load_game("connect_four")
batches = [game.new_batch(n) for _ in range(8)]
a32 = torch.randint(0, 7, (n,), dtype=torch.int32).pin_memory()
a8 = a32.to(torch.uint8).pin_memory()
mask = torch.empty((n, 1), dtype=torch.int32).pin_memory()
term = torch.empty((n,), dtype=torch.uint8).pin_memory()
rets = torch.empty((n, 2), dtype=torch.float32).pin_memory()
status = torch.empty((n,), dtype=torch.uint8).pin_memory()
out = {}
for name, call in (("compact", lambda b: b.step_host_compact(a8, status)), ("float", lambda b: b.step_host(a32, mask, term, rets))):
    for b in batches:
        b.reset(); call(b)
    best = 1e9
    for rep in range(4):
        for b in batches:
            b.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            for b in batches:
                call(b)
        best = min(best, (time.perf_counter() - t0) / 40)
    out[name] = best
print("zero_copy=%s (blocks/SM %s) graph=%s chunks=%s  compact %.1f us (%.3e steps/s)  float %.1f us (%.3e steps/s)  replays %d" % (
    os.environ.get("B2S_HOST_ZEROCOPY", "1"), os.environ.get("B2S_ZC_BLOCKS_PER_SM", "-"), os.environ.get("B2S_HOST_GRAPH", "1"), os.environ.get("B2S_HOST_CHUNKS", "default"), out["compact"] * 1e6, n / out["compact"],
    out["float"] * 1e6, n / out["float"], _lib.lib().b2s_host_graph_launches()) + "  zero-copy steps %d" % _lib.lib().b2s_host_zero_copy_steps())
'''
for zc, graph, chunks, per_sm in (("0", "0", None, None), ("0", "1", None, None), ("1", "1", None, "1"), ("1", "1", None, "2"), ("1", "1", None, "4"),
                                  ("1", "1", None, "8")):
    env = dict(os.environ, B2S_HOST_GRAPH=graph, B2S_HOST_ZEROCOPY=zc)
    if per_sm:
        env["B2S_ZC_BLOCKS_PER_SM"] = per_sm
    env.pop("B2S_HOST_CHUNKS", None)
    if chunks:
        env["B2S_HOST_CHUNKS"] = chunks
    subprocess.run([sys.executable, "-c", CODE], env=env, check=False)
