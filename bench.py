#!/usr/bin/env python3
"""bench.py — env steps/s of the batched connect_four ApplyAction hot path (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W          # our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N ...          # the CPU arm on the box's host cores

One "step" = one pass of State::ApplyAction over one batch of 1,048,576 connect_four states (SoA, 16 B per
state) with one legal action per state.  The (state, action) stream is synthetic: every lane is advanced
k_i ~ U{0..20} uniformly random legal plies from the start (non-terminal), then one uniformly random legal
action is drawn per lane.  Every timed step re-applies that action stream to a fresh
copy of the snapshot (the copy and an L2 flush happen outside the timed region), so all steps do equal work.

Printed JSON (one line, rank 0): metric/value = ApplyAction/s with states and actions resident in HBM;
e2e = the same step through b2s_step_fused_host with pinned HOST buffers (H2D actions, D2H mask/terminal/
returns inside the timed region); roofline = algorithmic bytes / CUDA-event time of the apply kernel vs the
measured HBM peak; cpu_baseline = the CPU arm on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STATES = 1 << 20
MAX_PREFIX = 20
BYTES_APPLY = 36          # 16 R state + 4 R action + 16 W state   
BYTES_FUSED = 49          # + 1 W terminal + 8 W returns + 4 W mask
METRIC = "connect_four ApplyAction env steps/sec (batched)"
UNIT = "steps/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ CPU arm

def cpu_arm(n_sample, threads, reps, seed=0x5EED):
    """ApplyAction of the CPU implementation on `threads` host threads over a bounded sample, `reps` passes.
    Returns (steps/s over all passes, kind, per-pass seconds).  kind = "reference" when oracle/_ref holds the
    unmodified reference build (oracle/ref_build.mk), else "port" (the oracle restatement)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_bench")
    if os.path.exists(ref):
        try:
            out = subprocess.run([ref, "apply", "connect_four", str(n_sample), str(MAX_PREFIX), str(seed),
                                  str(threads), str(reps)], capture_output=True, text=True, timeout=900)
            if out.returncode == 0:
                d = json.loads(out.stdout.strip().splitlines()[-1])
                return d["steps_per_s"], "reference", d["per_rep_seconds"]
        except Exception:
            pass
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import OracleGame, lib
    L = lib()
    L.orc_bench_apply.restype = C.c_double
    L.orc_bench_apply.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                  C.POINTER(C.c_double), C.POINTER(C.c_double)]
    g = OracleGame("connect_four")
    secs = C.c_double()
    per = (C.c_double * reps)()
    v = L.orc_bench_apply(g._g, n_sample, MAX_PREFIX, seed, threads, reps, C.byref(secs), per)
    return v, "port", list(per)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_sample = 1 << 18
    _, kind, per = cpu_arm(n_sample, cores, args.warmup + args.steps)
    times = per[args.warmup:]
    ms = 1e3 * sum(times) / max(len(times), 1)
    value = n_sample / (ms / 1e3) if ms > 0 else 0.0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "connect_four batched ApplyAction, SoA batch (CPU arm: one heap State per lane)",
                   "states_per_step": n_sample, "prefix_plies": "U{0..%d}" % MAX_PREFIX},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d states per step (same U{0..20}-ply synthetic stream), all host threads, Clone excluded" % n_sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm

class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0:
                    self.samples.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def build_workload(torch, game, n, dev, seed):
    """Snapshot batch of n non-terminal positions + one legal action per lane (all on device)."""
    snap = game.new_batch(n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    k = torch.randint(0, MAX_PREFIX + 1, (n,), device=dev, generator=gen)
    cols = torch.arange(7, device=dev, dtype=torch.int32)

    def random_legal(mask_words):
        legal = ((mask_words[:, :1] >> cols) & 1).bool()                       # [n,7]
        score = torch.rand((n, 7), device=dev, generator=gen).masked_fill(~legal, -1.0)
        return score.argmax(dim=1).to(torch.int32), legal.any(dim=1)

    # advance lane i by k_i plies, never stepping INTO a terminal state (keep the pre-terminal position)
    probe = game.new_batch(n)
    for t in range(MAX_PREFIX):
        a, has = random_legal(snap.legal_actions_mask_words())
        a = torch.where((t < k) & has, a, torch.full_like(a, -1))
        probe.copy_from(snap)
        probe.apply_actions(a)
        _, term, _ = probe.status()
        a = torch.where(term.bool(), torch.full_like(a, -1), a)               # do not enter terminal states
        snap.apply_actions(a)
    actions, has = random_legal(snap.legal_actions_mask_words())
    assert bool(has.all())
    snap.check_errors()
    return game, snap, actions.contiguous()


def run_gpu(args):
    import torch
    import open_spiel_b200 as b2
    from open_spiel_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = N_STATES                      # per GPU: weak scaling, independent shards, no data-path collective
    game = b2.Game("connect_four", device=local)
    _, snap, actions = build_workload(torch, game, n, dev, seed=0x5EED + rank)
    work = game.new_batch(n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    mask = torch.empty((n, 1), dtype=torch.int32, device=dev)
    term = torch.empty((n,), dtype=torch.uint8, device=dev)
    rets = torch.empty((n, 2), dtype=torch.float32, device=dev)
    # pinned host buffers for the end-to-end arm
    act_h = actions.cpu().pin_memory()
    mask_h = torch.empty((n, 1), dtype=torch.int32).pin_memory()
    term_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    rets_h = torch.empty((n, 2), dtype=torch.float32).pin_memory()

    L = _lib.lib()

    def prep():
        work.copy_from(snap)
        flush.fill_(rank + 1)           # evict the batch from L2 between timed iterations

    def timed(fn, iters, warm):
        evs = []
        for i in range(warm + iters):
            prep()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            if i >= warm:
                evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]      # ms per iteration (device time)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = L.b2s_launch_count()
    barrier()
    if sampler:
        sampler.start()
    # ---- headline: ApplyAction, device-resident ---------------------------------------------------
    t_apply = timed(lambda: work.apply_actions(actions), args.steps, args.warmup)
    launches_timed = args.steps          # one apply kernel per timed step (prep kernels are outside the events)
    barrier()
    # ---- extras: fused step, legal mask ---------------------------------------------------------------
    t_fused = timed(lambda: work.step(actions, mask, term, rets), args.steps, args.warmup)
    t_mask = timed(lambda: work.legal_actions_mask_words(out=mask), args.steps, args.warmup)
    barrier()
    work.check_errors()
    # ---- e2e: host buffers through b2s_step_fused_host (wall clock around a synchronous call) -----------
    e2e_times = []
    for i in range(args.warmup + args.steps):
        prep()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        work.step_host(act_h, mask_h, term_h, rets_h)
        t1 = time.perf_counter()
        if i >= args.warmup:
            e2e_times.append((t1 - t0) * 1e3)
    barrier()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    work.check_errors()
    total_launches = L.b2s_launch_count() - launches0

    def agg(ms_list):
        """max-over-ranks mean ms per step."""
        m = sum(ms_list) / len(ms_list)
        if dist is not None:
            t = torch.tensor([m], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            m = float(t.item())
        return m

    ms_apply, ms_fused, ms_mask, ms_e2e = agg(t_apply), agg(t_fused), agg(t_mask), agg(e2e_times)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak, peak_src = hbm_peak()
    value = world * n / (ms_apply / 1e3)
    ach = BYTES_APPLY * n / (ms_apply / 1e3) / 1e9          # per GPU
    h2d, d2h = 4 * n, (4 + 1 + 8) * n
    cores = os.cpu_count() or 1
    cpu_v, cpu_kind, cpu_per = cpu_arm(1 << 18, 1, 8)
    cpu_secs = sum(cpu_per)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_apply_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_apply, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": "connect_four batched ApplyAction, 1,048,576-state SoA batch per GPU (BASELINE configs[1])",
                   "states_per_step_per_gpu": n, "state_bytes": 16, "action_dtype": "int32",
                   "prefix_plies": "U{0..%d}" % MAX_PREFIX, "l2": "flushed between timed iterations (256 MiB write)",
                   "parallelism": "independent shards x%d, no data-path collective" % world},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "kernel": "k_apply<ConnectFourRules>", "bytes_per_step": BYTES_APPLY,
                     "peak_source": peak_src},
        "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                         "sample": "%d states x 8 passes, 1 thread, Clone excluded (%.2f s timed)" % (1 << 18, cpu_secs),
                         "host_cores": cores},
        "e2e": {"value": world * n / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e, "call": "b2s_step_fused_host (pinned host actions in; mask, terminal, returns out)"},
        "gpu_launches": launches_timed,
        "extras": {"fused_step_steps_per_s": world * n / (ms_fused / 1e3), "fused_ms": ms_fused,
                   "fused_gbs": BYTES_FUSED * n / (ms_fused / 1e3) / 1e9,
                   "legal_mask_per_s": world * n / (ms_mask / 1e3), "legal_mask_ms": ms_mask,
                   "launches_total_incl_setup": total_launches},
        "clocks": sampler.summary() if sampler else None,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
