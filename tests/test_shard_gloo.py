"""N > 1 host logic on CPU: two gloo ranks run the multi-rank plumbing bench.py uses on GPUs (dpose_b200/dist.py:
contiguous pose shards, barrier around the timed region, max-over-ranks time, whole-job throughput, the CPU-side
barrier of the one-process-all-GPUs leg) — one process per device, costmap replicated, no data-path collective.  The
results gathered on rank 0 must equal the single-process evaluation.  The compute stand-in is the CPU oracle (this
is a test; the product's compute path is CUDA-only and is covered by the -m gpu tests)."""
import os
import socket
import sys

import numpy as np
import pytest

import fixtures as fx

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dpose_b200.dist import Ranks, shard_bounds
    from oracle import capi
    ranks = Ranks(backend="gloo")
    assert (ranks.rank, ranks.world) == (rank, world)
    m = fx.bernoulli_map(300, 300, 0.10, 1)      # replicated costmap
    poses = fx.random_poses(1001, 300, 300, 2)   # strong scaling: every rank draws the batch, keeps its slice
    lo, hi = shard_bounds(len(poses), rank, world)
    oc = capi.CostData(fx.RECT, 2)
    ranks.barrier()
    part = oc.get_cost_batch(m, poses[lo:hi])
    ranks.barrier()
    seconds = 0.25 * (rank + 1)                  # a made-up per-rank duration: the job takes the slowest rank's
    tmax = ranks.max(seconds)
    total = ranks.sum(hi - lo)
    thr = ranks.throughput(hi - lo, seconds)
    ranks.host_barrier()
    gathered = [None] * world
    ranks.dist.all_gather_object(gathered, (lo, hi, part))
    if rank == 0:
        out = np.zeros((len(poses), 4))
        covered = np.zeros(len(poses), int)
        for lo_, hi_, p in gathered:
            out[lo_:hi_] = p
            covered[lo_:hi_] += 1
        np.savez(tmp, out=out, covered=covered, tmax=tmax, total=total, thr=thr)
    ranks.barrier()
    ranks.close()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import capi
    tmp = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    r = np.load(tmp)
    assert np.all(r["covered"] == 1)
    assert float(r["tmax"]) == 0.5 and float(r["total"]) == 1001 and float(r["thr"]) == 1001 / 0.5
    m = fx.bernoulli_map(300, 300, 0.10, 1)
    poses = fx.random_poses(1001, 300, 300, 2)
    assert np.array_equal(r["out"], capi.CostData(fx.RECT, 2).get_cost_batch(m, poses))


def _band_worker(rank, world, port, tmp):
    """strong scaling by map stripes: the job's poses ordered by map row once, contiguous slices, every rank evaluates
    its slice on ITS BAND of the replicated map only (the rows outside zeroed: what a rank that re-reads only its band
    of a volatile map sees)"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from dpose_b200.dist import Ranks, reach_of_box, row_band, row_order, shard_bounds
    from oracle import capi
    ranks = Ranks(backend="gloo")
    m = fx.bernoulli_map(300, 300, 0.10, 1)
    poses = fx.random_poses(1001, 300, 300, 2)
    poses = poses[row_order(poses[:, 1])]
    lo, hi = shard_bounds(len(poses), rank, world)
    oc = capi.CostData(fx.RECT, 2)
    band = row_band(poses[lo:hi, 1].min(), poses[lo:hi, 1].max(), reach_of_box(oc.box), 300)
    seen = np.zeros_like(m)
    seen[band[0]:band[1]] = m[band[0]:band[1]]
    part = oc.get_cost_batch(seen, poses[lo:hi])
    gathered = [None] * world
    ranks.dist.all_gather_object(gathered, (lo, hi, band, part))
    if rank == 0:
        out = np.zeros((len(poses), 4))
        for lo_, hi_, _, p in gathered:
            out[lo_:hi_] = p
        np.savez(tmp, out=out, bands=np.array([g[2] for g in gathered]))
    ranks.barrier()
    ranks.close()


def test_two_rank_row_bands_match_whole_map(tmp_path):
    import torch.multiprocessing as mp
    from dpose_b200.dist import row_order
    from oracle import capi
    tmp = str(tmp_path / "band.npz")
    mp.spawn(_band_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    r = np.load(tmp)
    m = fx.bernoulli_map(300, 300, 0.10, 1)
    poses = fx.random_poses(1001, 300, 300, 2)
    poses = poses[row_order(poses[:, 1])]
    assert np.array_equal(r["out"], capi.CostData(fx.RECT, 2).get_cost_batch(m, poses))  # every contribution lies in the band
    bands = r["bands"]
    assert all(0 <= b[0] < b[1] <= 300 and b[1] - b[0] < 200 for b in bands)  # ~ half the map + reach each, not the whole map
    assert bands[0][0] < bands[1][0] and bands[0][1] < bands[1][1] and bands[1][0] < bands[0][1]  # neighbours overlap by the reach


def test_row_band_and_reach():
    from dpose_b200.dist import reach_of_box, row_band, row_order
    from oracle import capi
    oc = capi.CostData(fx.RECT, 2)  # table 17 x 25, centre (12, 8): corners (-12, -8) .. (13, 9)
    assert reach_of_box(oc.box) == int(np.ceil(np.hypot(13, 9))) + 1
    assert row_band(100.2, 180.7, 17, 1000) == (83, 199)
    assert row_band(3.0, 5.0, 17, 1000) == (0, 23) and row_band(990.0, 999.5, 17, 1000) == (973, 1000)
    y = np.array([5.0, 1.0, 3.0, 1.0])
    assert row_order(y).tolist() == [1, 3, 2, 0]  # stable


def test_single_process_ranks_are_a_no_op():
    from dpose_b200.dist import Ranks
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    ranks = Ranks(backend="gloo")
    assert (ranks.rank, ranks.world, ranks.dist) == (0, 1, None)
    ranks.barrier()
    ranks.host_barrier()
    assert ranks.max(3.5) == 3.5 and ranks.sum(7) == 7 and ranks.throughput(10, 2.0) == 5.0
    ranks.close()


def test_shard_bounds_cover_exactly():
    from dpose_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 1001, 10_000_000):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
