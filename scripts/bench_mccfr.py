"""Throughput of the device external-sampling MCCFR vs the unmodified reference solver on the host CPU.
Usage: python scripts/bench_mccfr.py"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import open_spiel_b200 as b2


def dev(gs, K, iters):
    s = b2.ExternalSamplingMCCFRSolver(b2.load_game(gs), seed=1, traversals_per_update=K)
    s.run_iteration(2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run_iteration(iters)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"impl": "b200", "game": gs, "traversals_per_update": K, "iterations": iters, "seconds": round(dt, 4),
            "traversals_per_s": 2 * K * iters / dt, "iterations_per_s": iters / dt, "nash_conv": s.nash_conv()}


def ref(gs, iters):
    import ref_lib
    if not ref_lib.available():
        return None
    s = ref_lib.RefMCCFR(ref_lib.RefGame(gs), 1)
    s.iterate(50)
    t0 = time.perf_counter()
    s.iterate(iters)
    dt = time.perf_counter() - t0
    return {"impl": "reference", "game": gs, "cores": 1, "iterations": iters, "seconds": round(dt, 4),
            "traversals_per_s": 2 * iters / dt, "iterations_per_s": iters / dt, "nash_conv": s.nash_conv()}


if __name__ == "__main__":
    for gs in ("leduc_poker", "kuhn_poker"):
        r = ref(gs, 20000)
        if r:
            print(json.dumps(r), flush=True)
        for K, iters in [(1, 2000), (256, 500), (4096, 200), (16384, 100), (65536, 30)]:
            print(json.dumps(dev(gs, K, iters)), flush=True)
