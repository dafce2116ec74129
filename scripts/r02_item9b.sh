# Second pass of scripts/r02_item9.sh after the coalesced mask stores and the independent prefetch loads of k_mccfr_partial_log
# ("after2"; the "before" files of the first pass stay valid).
cd /root/repo
mkdir -p gpurun_out
echo "== correctness of what changed"
timeout 1200 python -m pytest tests/test_gpu_mccfr.py tests/test_gpu_parity_games.py tests/test_gpu_vs_reference.py tests/test_gpu_trajectories.py tests/test_gpu_bench_workload.py -x -q -m gpu 2>&1 | tail -4
echo "== sweep after2"
timeout 600 python scripts/sweep_games.py 22 > gpurun_out/r02_sweep_games_after2.jsonl 2> gpurun_out/r02_sweep_after2.err
python - <<'P'
import json
for d in map(json.loads, open("gpurun_out/r02_sweep_games_after2.jsonl")):
    if d["kernel"] in ("legal_mask", "step_fused"):
        print("%-18s %-12s %.4f ms  %.3f of peak  (%d lanes)" % (d["game"], d["kernel"], d["ms"], d["frac_of_peak"], d["lanes"]))
P
echo "== MCCFR: scatter (default) / lanes / dense"
for mode in scatter lanes dense; do
  B2S_MCCFR_MODE=$mode timeout 600 python scripts/bench_mccfr.py > gpurun_out/r02_mccfr_$mode.jsonl 2> gpurun_out/r02_mccfr_$mode.err
  echo "-- $mode"; grep b200 gpurun_out/r02_mccfr_$mode.jsonl | grep leduc | cut -c1-200
done
echo "== ncu"
NCU="ncu --set full --clock-control none --import-source on -c 2"
# (bt mask capture already taken in the previous pass)
cat > /tmp/mc1.py <<'P'
import sys
sys.path.insert(0, ".")
import torch
import open_spiel_b200 as b2
s = b2.ExternalSamplingMCCFRSolver(b2.load_game("leduc_poker"), seed=1, traversals_per_update=16384)
s.run_iteration(3)
torch.cuda.synchronize()
P
ncu --set full --clock-control none --import-source on -k regex:"k_mccfr_(scatter|apply|es)" -s 6 -c 3 -f -o gpurun_out/r02_prof_mccfr_scatter python /tmp/mc1.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
