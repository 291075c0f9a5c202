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
