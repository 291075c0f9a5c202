"""Pose-batch sharding: contiguous slices, costmap and table replicated,
no data-path collective.  Same split as dpose_pg_get_cost_batch_multi (csrc/api.cu)."""


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's contiguous slice of n poses over `world` devices"""
    per = (n + world - 1) // world
    return min(n, per * rank), min(n, per * (rank + 1))
