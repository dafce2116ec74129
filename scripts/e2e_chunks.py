"""e2e host-buffer step time vs B2S_HOST_CHUNKS (run each value in a fresh process)."""
import os
import subprocess
import sys

CODE = r'''
import sys, time, torch
sys.path.insert(0, ".")
import open_spiel_b200 as b2
n = 1 << 20
game = b2.load_game("connect_four")
batches = [game.new_batch(n) for _ in range(8)]
acts = [torch.randint(0, 7, (n,), dtype=torch.int32).pin_memory() for _ in range(8)]
mask = torch.empty((n, 1), dtype=torch.int32).pin_memory()
term = torch.empty((n,), dtype=torch.uint8).pin_memory()
rets = torch.empty((n, 2), dtype=torch.float32).pin_memory()
for b in batches:
    b.step_host(acts[0], mask, term, rets, n=0)
for rep in range(3):
    for b in batches:
        b.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5):
        for j, b in enumerate(batches):
            b.step_host(acts[(i + j) % 8], mask, term, rets)
    dt = (time.perf_counter() - t0) / 40
print("chunks=%s ms_per_step=%.4f steps_per_s=%.3e" % (__import__("os").environ.get("B2S_HOST_CHUNKS"), dt * 1e3, n / dt))
'''
for c in (1, 2, 3, 4, 8):
    env = dict(os.environ, B2S_HOST_CHUNKS=str(c))
    subprocess.run([sys.executable, "-c", CODE], env=env, check=False)
