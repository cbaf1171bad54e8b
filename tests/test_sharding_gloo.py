"""N>1 host path on CPU: two gloo ranks shard one corpus by file, index their shards with the real host index pass
(wvb_index_many) and agree on totals; the timing reduction is a MAX over ranks.  No data-path collective is involved."""
import os
import socket
import sys

import numpy as np
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
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _harness import make_file
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import Corpus
    from wavpackdecoder_b200.sharding import max_over_ranks, shard_range, sum_over_ranks
    files = [make_file(seed=100 + i, seconds=0.2 + 0.05 * (i % 3), channels=1 + (i % 2))[2] for i in range(7)]
    lo, hi = shard_range(len(files), rank, world)
    cp = Corpus.from_files(files[lo:hi], out_format=N.OUT_PCM, threads=2)
    total_blocks = sum_over_ranks(cp.nblocks)
    total_samples = sum_over_ranks(cp.total_samples)
    t = max_over_ranks(1.0 + rank)
    whole = Corpus.from_files(files, out_format=N.OUT_PCM, threads=2)
    q.put((rank, lo, hi, total_blocks, total_samples, t, whole.nblocks, whole.total_samples))
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce():
    from wavpackdecoder_b200 import build
    build.build()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, tb0, ts0, t0, wb, ws), (r1, lo1, hi1, tb1, ts1, t1, _, _) = res
    assert (lo0, hi0, lo1, hi1) == (0, 4, 4, 7)
    assert tb0 == tb1 == wb and ts0 == ts1 == ws
    assert t0 == t1 == 2.0


def test_shard_helpers():
    from wavpackdecoder_b200.sharding import shard_by_cost, shard_range
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    parts = shard_by_cost([5, 1, 1, 1, 4, 3], 2)
    assert sorted(np.concatenate(parts).tolist()) == list(range(6))
    loads = [sum([5, 1, 1, 1, 4, 3][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 1
