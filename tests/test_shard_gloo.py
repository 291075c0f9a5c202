"""N>1 host logic on CPU: two gloo ranks shard a pose batch exactly like the multi-GPU driver
(one process per device, costmap replicated, no data-path collective), results gathered on rank 0
equal the single-process evaluation, and the timing reduction is a MAX over ranks.  The compute
stand-in is the CPU oracle (this is a test; the product path is CUDA-only)."""
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
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dpose_b200.shard import shard_bounds
    from oracle import capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = fx.bernoulli_map(300, 300, 0.10, 1)      # replicated costmap
    poses = fx.random_poses(1001, 300, 300, 2)   # every rank knows the batch, evaluates its slice
    lo, hi = shard_bounds(len(poses), rank, world)
    oc = capi.CostData(fx.RECT, 2)
    part = oc.get_cost_batch(m, poses[lo:hi])
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, part))
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)     # bench.py's max-over-ranks timing reduction
    if rank == 0:
        out = np.zeros((len(poses), 4))
        covered = np.zeros(len(poses), int)
        for lo_, hi_, p in gathered:
            out[lo_:hi_] = p
            covered[lo_:hi_] += 1
        np.savez(tmp, out=out, covered=covered, tmax=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import capi
    tmp = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    r = np.load(tmp)
    assert np.all(r["covered"] == 1)
    assert r["tmax"][0] == 2.0
    m = fx.bernoulli_map(300, 300, 0.10, 1)
    poses = fx.random_poses(1001, 300, 300, 2)
    assert np.array_equal(r["out"], capi.CostData(fx.RECT, 2).get_cost_batch(m, poses))


def test_shard_bounds_cover_exactly():
    from dpose_b200.shard import shard_bounds
    for n in (0, 1, 7, 8, 1001, 10_000_000):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
