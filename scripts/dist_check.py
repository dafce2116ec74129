#!/usr/bin/env python3
"""Run under torchrun (nccl), one rank per GPU: checks that sharded results are GPU-count invariant.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import open_spiel_b200 as b2  # noqa: E402
from open_spiel_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

# 1. CFR: traversal work split over ranks, NCCL all-reduce of the delta buffer
game = b2.Game("leduc_poker", device=local)
d = parallel.DistributedCFRSolver(game)
d.evaluate_and_update_policy(40)
single = b2.CFRSolver(game)
single.evaluate_and_update_policy(40)
td, ts = d.table(), single.table()
err = max(float(np.abs(td[f] - ts[f]).max()) for f in ("regrets", "cum_policy", "cur_policy"))
assert err == 0.0, err          # the sharded exchange is exact (one rank's value + zeros per slot)

# 1b. external-sampling MCCFR: reduction lanes dealt out to the ranks, NCCL all-gather -> bit-identical tables
import time  # noqa: E402
K = 16384
dm = parallel.DistributedExternalSamplingMCCFRSolver(game, seed=5, traversals_per_update=K)
sm = b2.ExternalSamplingMCCFRSolver(game, seed=5, traversals_per_update=K)
dm.run_iteration(6)
sm.run_iteration(6)
tdm, tsm = dm.table(), sm.table()
assert all(np.array_equal(tdm[f], tsm[f]) for f in ("regrets", "cum_policy")), "sharded MCCFR tables differ"
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter(); dm.run_iteration(40); torch.cuda.synchronize(); dist.barrier(); t_d = time.perf_counter() - t0
t0 = time.perf_counter(); sm.run_iteration(40); torch.cuda.synchronize(); t_s = time.perf_counter() - t0
mccfr_note = "mccfr K=%d bit-identical; %d-GPU %.3e trav/s vs 1-GPU %.3e" % (K, world, 2 * K * 40 / t_d, 2 * K * 40 / t_s)

# 2. MCTS: trees sharded by root index; per-tree results must not depend on the number of GPUs
g2 = b2.Game("connect_four", device=local)
total = 4096
lo, hi = parallel.shard_range(total)
mine = b2.mcts_search(g2.new_batch(hi - lo), 64, solve=False, seed=11, tree_index_offset=lo)
full = b2.mcts_search(g2.new_batch(total), 64, solve=False, seed=11)
assert torch.equal(mine["visits"], full["visits"][lo:hi]) and torch.equal(mine["total_reward"], full["total_reward"][lo:hi])

# 3. rollouts: lane_offset = global lane id; statistics all-reduced
g3 = b2.Game("breakthrough", device=local)
n = 1 << 16
lo, hi = parallel.shard_range(n)
b = g3.new_batch(hi - lo)
rets, plies = b.rollout(seed=3, lane_offset=lo)
stats = parallel.rollout_stats(rets, plies)
if rank == 0:
    bf = g3.new_batch(n)
    r2, p2 = bf.rollout(seed=3)
    s2 = parallel.rollout_stats.__wrapped__(r2, p2) if hasattr(parallel.rollout_stats, "__wrapped__") else None
    r0 = r2[:, 0]
    want = [int((r0 > 0).sum()), int((r0 < 0).sum()), int((r0 == 0).sum()), int(p2.sum()), n]
    assert stats.tolist() == want, (stats.tolist(), want)
    print("dist_check ok: world=%d cfr_max_abs_diff=%.3e rollout_stats=%s; %s" % (world, err, stats.tolist(), mccfr_note))
dist.barrier()
dist.destroy_process_group()
