"""Rank partitioning used by bench.py and the multi-GPU tests.

Two natural shards exist on this path:
  * independent frame pairs / camera streams -> contiguous blocks per rank, NO collective in the data path
    (weak scaling; only the timing is max-reduced)
  * one huge pair -> contiguous slices of the ordered point list per rank, one 232-byte all-reduce of the normal
    equations per evaluation (ea_shard_solve, NCCL)
"""


def shard_range(n, rank, world):
    """[begin, end) of rank's contiguous share of n ordered units (same rule as the device code: n*r/w)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_pairs(ref_idx, now_idx, rank, world):
    b, e = shard_range(len(ref_idx), rank, world)
    return ref_idx[b:e], now_idx[b:e]


def reduce_max(value, dist=None, device=None):
    """Max over ranks of a python float (timings are reported as the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value, dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, ms_this_rank, dist=None, device=None):
    """Whole-job units/s: all ranks' units divided by the slowest rank's time."""
    total = reduce_sum(units_this_rank, dist, device)
    ms = reduce_max(ms_this_rank, dist, device)
    return total / (ms * 1e-3), ms
