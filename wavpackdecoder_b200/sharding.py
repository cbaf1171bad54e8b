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


def shard_contiguous_by_cost(costs, world):
    """Contiguous [lo, hi) file ranges whose cost sums are as equal as a cut between files allows (cuts at the cost
    quantiles).  Contiguous shards keep a rank's files adjacent in the compressed slab, so its upload is one range."""
    c = np.cumsum(np.asarray(costs, dtype=np.float64))
    n = len(c)
    total = float(c[-1]) if n else 0.0
    cuts = [0]
    for r in range(1, world):
        k = int(np.searchsorted(c, total * r / world, side="left")) + 1 if total > 0 else (n * r) // world
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def file_costs(corpus):
    """Decode cost per file of an indexed Corpus: sum over its blocks of block_samples x (decorrelation passes + the
    entropy decoder, weighted like three passes) x coded channels.  DSD blocks count their byte-times."""
    from . import _native as N
    t = N.desc_table(corpus.descs, corpus.nblocks)
    stereo = ((t["flags"] & (4 | 0x40000000)) == 0).astype(np.float64) + 1.0
    per_block = t["block_samples"].astype(np.float64) * (t["sub_len"][:, N.SUB_TERMS].astype(np.float64) + 3.0) * stereo
    first = np.asarray(corpus.first, dtype=np.int64)
    count = np.asarray(corpus.count, dtype=np.int64)
    csum = np.concatenate([[0.0], np.cumsum(per_block)])
    return csum[first + count] - csum[first]


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
