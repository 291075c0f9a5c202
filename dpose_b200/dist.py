"""Host-side plumbing of the multi-GPU runs: one process per GPU, pose batches sharded contiguously,
costmap and table replicated, NO collective on the data path.  torch.distributed is used only for what surrounds the
hot path: the barrier around a timed region, the max-over-ranks reduction of its duration and the gathering of
per-rank counters.  The same functions run under NCCL (bench.py on GPUs) and under gloo (tests/test_shard_gloo.py on
CPU), the tensor device being the only difference.
"""
import os


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous slice of n poses over `world` devices — the split of
    dpose_pg_get_cost_batch_multi (csrc/api.cu)"""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))


def row_order(y):
    """order of a pose set by map row (stable argsort of the y coordinates): what a caller does once, at set-up, so that
    the contiguous slices of shard_bounds become stripes of the map ("pre-sorted by map tile so each
    GPU's slice touches a compact region")"""
    import numpy as np
    return np.argsort(np.asarray(y, dtype=np.float64), kind="stable")


def reach_of_box(box):
    """dpose_pg_reach from cost_data::get_box() (5 x 2 ring around the centre): the farthest corner of the cost table from
    its centre, rounded up, + 1 — no cell farther than this from a pose contributes to its cost"""
    import math
    import numpy as np
    b = np.asarray(box, dtype=np.float64).reshape(-1, 2)[:4]
    return int(math.ceil(math.sqrt(float((b * b).sum(axis=1).max())))) + 1


def row_band(y_min, y_max, reach, size_y):
    """[row_lo, row_hi) of the map rows that can contribute to poses with y in [y_min, y_max]: the stripe grown by the
    footprint's reach (pose_gradient.reach()), clamped to the map — the band a rank hands to
    dpose_pg_get_cost_batch_device_rows with its slice of a row-ordered pose set"""
    import math
    lo = max(0, int(math.floor(y_min)) - int(reach))
    hi = min(int(size_y), int(math.ceil(y_max)) + int(reach) + 1)
    return lo, max(hi, lo + 1)


class Ranks:
    """rank / world of this process and the reductions a timed multi-rank run needs"""

    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.dist = None
        self.host_group = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend == "nccl":
                dist.init_process_group(backend="nccl", device_id=device)
                # waits that must leave every GPU idle (rank 0 driving all devices from one process) go through a
                # CPU-side group: an NCCL barrier would park a spinning kernel on each GPU
                self.host_group = dist.new_group(backend="gloo")
            else:
                dist.init_process_group(backend=backend or "gloo", rank=self.rank, world_size=self.world)
            self.dist = dist

    def _tensor(self, values):
        import torch
        return torch.tensor(values, dtype=torch.float64, device=self.device if self.device is not None else "cpu")

    def barrier(self):
        """all ranks have reached this point and (on GPUs) their queued work is finished"""
        import torch
        if self.dist is not None:
            self.dist.barrier()
        if self.device is not None and torch.cuda.is_available():
            torch.cuda.synchronize()

    def host_barrier(self):
        """CPU-side barrier: no kernel is parked on any GPU while ranks wait"""
        if self.dist is not None:
            self.dist.barrier(group=self.host_group) if self.host_group is not None else self.dist.barrier()

    def max(self, value):
        """max over ranks of a per-rank duration: the time of the whole job"""
        t = self._tensor([float(value)])
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, value):
        t = self._tensor([float(value)])
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def throughput(self, units_this_rank, seconds_this_rank):
        """whole-job throughput: units all ranks processed / the slowest rank's time"""
        return self.sum(units_this_rank) / self.max(seconds_this_rank)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
