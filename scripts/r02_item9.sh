# Before / after measurements of the three lower-priority kernels (round-2 item 9), one GPU:
#   breakthrough legal mask (piece loop + indexed local array  ->  bit-parallel spread),
#   leduc_poker streaming kernels (unpack / pack of 17 fields  ->  in-place bit-field edits),
#   MCCFR table update (dense [K][E] delta rows  ->  delta logs).
# "before" = scripts/micro/_before/libb2s.so, the library built from commit 8772d28 (same C ABI), selected with B2S_LIBRARY;
# the MCCFR "before" is the same library as "after" with B2S_MCCFR_DENSE=1.
cd /root/repo
mkdir -p gpurun_out
BEFORE=/root/repo/scripts/micro/_before/libb2s.so
echo "== correctness of what changed"
timeout 900 python -m pytest tests/test_gpu_mccfr.py tests/test_gpu_parity_games.py tests/test_gpu_vs_reference.py tests/test_gpu_mcts.py -x -q -m gpu -k "mccfr or leduc or breakthrough or delta" 2>&1 | tail -4
echo "== sweep (4M lanes; 256k for wide masks), after / before"
timeout 600 python scripts/sweep_games.py 22 > gpurun_out/r02_sweep_games_after.jsonl 2> gpurun_out/r02_sweep_after.err
B2S_LIBRARY=$BEFORE timeout 600 python scripts/sweep_games.py 22 > gpurun_out/r02_sweep_games_before.jsonl 2> gpurun_out/r02_sweep_before.err
python - <<'P'
import json
def load(f):
    return {(d["game"], d["kernel"]): d for d in map(json.loads, open(f))}
a, b = load("gpurun_out/r02_sweep_games_after.jsonl"), load("gpurun_out/r02_sweep_games_before.jsonl")
for k in a:
    if k[0] in ("breakthrough", "leduc_poker"):
        print("%-14s %-12s before %.3f ms (%.3f of peak) -> after %.3f ms (%.3f of peak)" % (k[0], k[1], b[k]["ms"], b[k]["frac_of_peak"], a[k]["ms"], a[k]["frac_of_peak"]))
P
echo "== MCCFR: delta logs vs dense rows"
timeout 600 python scripts/bench_mccfr.py > gpurun_out/r02_mccfr_log.jsonl 2> gpurun_out/r02_mccfr_log.err
B2S_MCCFR_DENSE=1 timeout 600 python scripts/bench_mccfr.py > gpurun_out/r02_mccfr_dense.jsonl 2> gpurun_out/r02_mccfr_dense.err
grep b200 gpurun_out/r02_mccfr_log.jsonl | cut -c1-220
grep b200 gpurun_out/r02_mccfr_dense.jsonl | cut -c1-220
echo "== ncu"
NCU="ncu --set full --clock-control none --import-source on -c 2"
$NCU -k regex:k_legal_mask -o gpurun_out/r02_prof_bt_mask_after python scripts/profile_one.py legal_mask 18 breakthrough > /dev/null 2>&1
B2S_LIBRARY=$BEFORE $NCU -k regex:k_legal_mask -o gpurun_out/r02_prof_bt_mask_before python scripts/profile_one.py legal_mask 18 breakthrough > /dev/null 2>&1
$NCU -k regex:k_apply -o gpurun_out/r02_prof_leduc_apply_after python scripts/profile_one.py apply 22 leduc_poker > /dev/null 2>&1
B2S_LIBRARY=$BEFORE $NCU -k regex:k_apply -o gpurun_out/r02_prof_leduc_apply_before python scripts/profile_one.py apply 22 leduc_poker > /dev/null 2>&1
cat > /tmp/mc1.py <<'P'
import sys
sys.path.insert(0, ".")
import torch
import open_spiel_b200 as b2
s = b2.ExternalSamplingMCCFRSolver(b2.load_game("leduc_poker"), seed=1, traversals_per_update=16384)
s.run_iteration(3)
torch.cuda.synchronize()
P
ncu --set full --clock-control none --import-source on -k regex:"k_mccfr_(partial_log|combine|es)" -s 6 -c 3 -o gpurun_out/r02_prof_mccfr_log python /tmp/mc1.py > /dev/null 2>&1
B2S_MCCFR_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:"k_mccfr_(apply|es)" -s 4 -c 2 -o gpurun_out/r02_prof_mccfr_dense python /tmp/mc1.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== bench at the driver's K (20 steps, 5 warm-up)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r02_bench_k20.err | tail -1 > gpurun_out/r02_bench_k20.json
python - <<'P'
import json
d = json.load(open("gpurun_out/r02_bench_k20.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_events_around_graph"), d["config"].get("events_in_graph"), d["e2e"]["value"])
P
tail -2 gpurun_out/r02_bench_k20.err
