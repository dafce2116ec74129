#!/usr/bin/env python3
"""Turn gpurun_out/{launches.csv,prof_*.ncu-rep,bench*.json,sweep.jsonl} into tracked summaries under profiles/.
Usage: python scripts/summarize_profiles.py r01"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
           "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]

# 1. launch list
lp = os.path.join(GO, tag + "_launches.csv")
if not os.path.exists(lp):
    lp = os.path.join(GO, "launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi:
            try:
                d[r[ki]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(OUT, tag + "_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ python bench.py (see scripts/gpu_check.sh for the loop sizes of the profiled run)\n")
        f.write("# per-launch times are cold-cache and serialised: compare shares, not absolutes.\n")
        f.write("# the headline timed region contains ONLY k_apply launches (K of them in one CUDA graph);\n")
        f.write("# k_copy = untimed restore of the per-step batches, k_step_fused / k_legal_mask = the 'extras' timings.\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-110s n=%5d mean_ns=%10.1f share=%5.1f%%\n" % (k[:110], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    print("wrote launches summary")

# 2. full captures
for rep in sorted(os.listdir(GO)):
    if not rep.endswith(".ncu-rep") or (rep.startswith("r0") and not rep.startswith(tag + "_")):
        continue
    name = rep[:-8]
    if name.startswith(tag + "_"):
        name = name[len(tag) + 1:]
    raw = subprocess.run(["ncu", "-i", os.path.join(GO, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        rec = {"kernel": row[hdr.index("Kernel Name")]}
        for m in METRICS:
            if m in hdr:
                rec[m] = row[hdr.index(m)] + " " + units[hdr.index(m)]
        out.append(rec)
    with open(os.path.join(OUT, "%s_%s_ncu_raw.json" % (tag, name)), "w") as f:
        json.dump(out, f, indent=1)
    if name == "prof_apply":
        def num(rec, key):
            v, u = rec[key].split(" ")[0], rec[key].split(" ")[1] if " " in rec[key] else ""
            x = float(v.replace(",", ""))
            return x * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
        tr = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in out]
        with open(os.path.join(OUT, tag + "_apply_traffic.json"), "w") as f:
            json.dump({"dram_bytes_per_launch": sum(tr) / len(tr),
                       "note": "dram__bytes_read.sum + dram__bytes_write.sum per k_apply launch (ncu --set full, 1M states). "
                               "ncu invalidates caches before each replay and the 16.8 MB of written state stays dirty in L2 until "
                               "evicted after the kernel, so the write half of the 37.7 MB algorithmic traffic is not inside the window.",
                       "algorithmic_bytes_per_launch": 36 * (1 << 20)}, f, indent=1)
    print("wrote", name)

# 3. bench lines / sweep
for fn in ("bench.json", "bench_ref.json", "gpu.csv", "nproc.txt", "pytest_gpu.log", "smoke.log"):
    p = os.path.join(GO, tag + "_" + fn)
    if not os.path.exists(p):
        p = os.path.join(GO, fn)
    if os.path.exists(p):
        with open(os.path.join(OUT, tag + "_" + fn), "w") as f:
            f.write(open(p).read())
