"""Multi-GPU plumbing: shards share nothing, so this is only file partitioning plus the
barrier / max-over-ranks reductions a measurement needs.  No data-path collective exists."""
import numpy as np


def shard_range(nfiles, rank, world):
    """Contiguous, balanced [lo, hi) of files for this rank (sizes differ by at most one)."""
    base, extra = divmod(int(nfiles), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs, world):
    """Greedy longest-first partition of files by a cost (e.g. sum of block_samples x terms) for unequal files.
    Returns a list of index arrays, one per rank."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    load = np.zeros(world)
    parts = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(load))
        parts[r].append(int(i))
        load[r] += float(costs[i])
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float over the default process group (identity without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
