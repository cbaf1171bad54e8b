#!/usr/bin/env python3
"""Small device-resident decode loop for profiling (ncu) and quick kernel timing.
    python tools/prof_run.py --files 2000 --seconds 1 --steps 2 [--kw bits=24 ...]
Prints kernel ms per step and samples/s."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=2000)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--fmt", default="pcm")
    ap.add_argument("--kw", nargs="*", default=[], help="encoder config overrides, e.g. bits=24 channels=6 terms=18,18,2 kind=1")
    ap.add_argument("--open-flags", type=lambda x: int(x, 0), default=0)
    args = ap.parse_args()
    import torch
    import bench
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    kw = {}
    for item in args.kw:
        k, v = item.split("=", 1)
        if k in ("terms", "deltas"):
            kw[k] = [int(x) for x in v.split(",") if x]
        else:
            kw[k] = int(v, 0)
    corpus = bench.build_corpus(args.files, args.seconds, 0x5EED0000, bench.host_cores(), 30.0, pin=False, cfg_kw=kw)
    slab = corpus["slab"]
    fmt = N.OUT_PCM if args.fmt == "pcm" else N.OUT_INT32
    cp = Corpus(slab, corpus["offsets"], corpus["sizes"], out_format=fmt, open_flags=args.open_flags)
    dec = BatchDecoder(0)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(slab).to(dev)
    d_out = torch.empty(cp.out_bytes + 64, dtype=torch.uint8, device=dev)
    d_res = torch.empty(max(cp.nblocks, 1) * 16, dtype=torch.uint8, device=dev)
    dec.prepare(cp.descs, cp.nblocks, fmt)
    FL = N.IN_DEVICE | N.OUT_DEVICE | N.RESULTS_DEVICE
    for i in range(args.steps):
        dec.decode(d_in.data_ptr(), slab.size, None, cp.nblocks, d_out.data_ptr(), cp.out_bytes, fmt, FL, d_res.data_ptr())
        tm = dec.timing()
        print("step %d: kernel %.3f ms, %d blocks, %.3f Gsamples/s, launches %d, in %.2f GB out %.2f GB -> %.1f GB/s" % (
            i, tm["kernel_ms"], cp.nblocks, cp.total_samples / tm["kernel_ms"] / 1e6, tm["launches"], corpus["compressed_bytes"] / 1e9,
            cp.out_bytes / 1e9, (corpus["compressed_bytes"] + cp.out_bytes) / tm["kernel_ms"] / 1e6), flush=True)
    res = d_res.cpu().numpy().view(np.uint32).reshape(-1, 4)
    print("flagged blocks:", int((res[:cp.nblocks, 1] != 0).sum()))


if __name__ == "__main__":
    main()
