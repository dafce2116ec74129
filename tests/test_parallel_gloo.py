"""CPU, world_size 2, gloo: the N>1 host logic — lane sharding, random-stream offsets, statistics all-reduce."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from open_spiel_b200 import parallel
    assert parallel.world() == (rank, world)
    total = 1_000_003
    lo, hi = parallel.shard_range(total)
    # every rank's slice, gathered: contiguous, disjoint, covering
    t = torch.tensor([lo, hi], dtype=torch.int64)
    got = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(got, t)
    edges = [g.tolist() for g in got]
    assert edges[0][0] == 0 and edges[-1][1] == total
    for a, b in zip(edges, edges[1:]):
        assert a[1] == b[0]
    assert max(e[1] - e[0] for e in edges) - min(e[1] - e[0] for e in edges) <= 1
    # statistics all-reduce: finished playouts of this rank's lanes
    n = hi - lo
    gen = torch.Generator().manual_seed(rank)
    r0 = torch.randint(-1, 2, (n,), generator=gen).float()
    returns = torch.stack([r0, -r0], dim=1)
    plies = torch.randint(7, 43, (n,), generator=gen)
    stats = parallel.rollout_stats(returns, plies)
    local = torch.tensor([(r0 > 0).sum(), (r0 < 0).sum(), (r0 == 0).sum(), plies.sum(), n])
    both = [torch.zeros(5, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(both, local.to(torch.int64))
    assert torch.equal(stats, both[0] + both[1])
    assert int(stats[4]) == total and int(stats[:3].sum()) == total
    # lane-sharded MCCFR exchange: every rank owns 64 / world reduction lanes; after the gather all ranks hold the same
    # [64, E] partial sums, so the fixed-order tree gives every rank the same bits
    llo, lhi = parallel.lane_range()
    assert (llo, lhi) == (rank * 32, rank * 32 + 32)
    E = 37
    full = torch.arange(64 * E, dtype=torch.float64).reshape(64, E) * 0.1 + 1.0 / 3.0
    partials = torch.zeros(64, E, dtype=torch.float64)
    partials[llo:lhi] = full[llo:lhi]
    parallel.gather_lanes(partials, llo, lhi)
    assert torch.equal(partials, full)
    q.put((rank, edges, stats.tolist()))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_stats():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = [q.get(timeout=5) for _ in range(world)]
    assert results[0][1] == results[1][1] and results[0][2] == results[1][2]     # all ranks agree


def test_lane_range_properties():
    sys.path.insert(0, ROOT)
    from open_spiel_b200.parallel import lane_range
    for world in (1, 2, 4, 8, 16, 32, 64):
        pieces = [lane_range(r, world) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == 64 and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    with pytest.raises(ValueError):
        lane_range(0, 3)


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from open_spiel_b200.parallel import shard_range
    for total in (0, 1, 7, 8, 1 << 20, 1_000_003):
        for world in (1, 2, 3, 4, 8):
            pieces = [shard_range(total, r, world) for r in range(world)]
            assert pieces[0][0] == 0 and pieces[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
