#!/usr/bin/env python3
"""Fuzz the index pass + device decode function (host emulation, tests/emul) against the oracle on randomly damaged streams.
A mismatch is only acceptable when a block result carries WVB_RF_INEXACT (documented corners) or the oracle itself raised.
    python tools/fuzz_parity.py [iterations] [seed]"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _harness import KIND_DSD, KIND_FLOAT, KIND_HYBRID, emul_decode_file, make_file, oracle_decode  # noqa: E402

BASES = [dict(), dict(channels=1), dict(bits=24), dict(kind=KIND_HYBRID), dict(bits=32, int32_sent_bits=8), dict(kind=KIND_FLOAT, bits=32),
         dict(terms=[17, 3, -1, 8]), dict(false_stereo=1), dict(kind=KIND_DSD, dsd_mode=1, block_samples=6000),
         dict(kind=KIND_DSD, dsd_mode=3, block_samples=6000), dict(kind=KIND_DSD, dsd_mode=0, block_samples=6000), dict(extras=1 | 2 | 4 | 8 | 16 | 64),
         dict(terms=[18, 18, 2, 17, 3]), dict(terms=[18, 17]), dict(bits=8, channels=1), dict(bits=24, channels=1, block_samples=1001),
         # round 2: the 16-term in-register list, the stock lists under hybrid / float / int32 (in-register fixup kernels), mono
         # blocks that start inside a caller's call (quirk C-5), block checksums
         dict(terms=[18, 18, 2, 3, -2, 18, 2, 4, 7, 5, 3, 6, 8, -1, 18, 2], deltas=[2] * 16), dict(kind=KIND_HYBRID, channels=1, terms=[18, 18, 2, 3], deltas=[2] * 4),
         dict(kind=KIND_FLOAT, bits=32, channels=1, terms=[18, 18, 2, 3], deltas=[2] * 4), dict(channels=1, block_samples=1000, terms=[18, 18, 2, 3], deltas=[2] * 4),
         dict(channels=1, block_samples=700, bits=24), dict(extras=2 | 8, block_samples=3000)]


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    files = []
    for i, kw in enumerate(BASES):
        kw = dict(kw)
        secs = 0.05 if kw.get("kind") == KIND_DSD else 0.4
        kw.setdefault("block_samples", 5000)
        files.append(make_file(seed=900 + i, seconds=secs, **kw)[2])
    stats = dict(ok=0, inexact=0, oracle_exc=0, open_err=0, mismatch=0)
    for it in range(iters):
        data = bytearray(rng.choice(files))
        kind = rng.random()
        if kind < 0.6:
            for _ in range(rng.choice([1, 1, 2, 5])):
                data[rng.randrange(len(data))] ^= 1 << rng.randrange(8)
        elif kind < 0.8:
            data = data[: rng.randrange(40, len(data))]
        elif kind < 0.9:
            a = rng.randrange(len(data)); b = min(len(data), a + rng.randrange(1, 3000))
            del data[a:b]
        else:
            a = rng.randrange(len(data))
            data[a:a] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 200)))
        data = bytes(data)
        chunk = rng.choice([4096, 4096, 1000, 333])
        # a flipped bit in total_samples makes the reference pad zeros up to that count (billions of samples): not worth the minutes
        tot = max((int.from_bytes(data[p + 12:p + 16], "little") for p in range(0, max(0, len(data) - 32)) if data[p:p + 4] == b"wvpk"), default=0)
        # ... and so does a flipped bit in a block_index (a gap of billions of samples): ask the index pass how much would come out
        huge = tot != 0xffffffff and tot > 4000000
        if not huge:
            try:
                from _harness import emul, index_file
                finfo0, _d, _n = index_file(emul(), data, 0, chunk)
                huge = finfo0.status == 0 and finfo0.indexed_samples > 4000000
            except Exception:
                pass
        if huge:
            stats["skipped_huge"] = stats.get("skipped_huge", 0) + 1
            continue
        try:
            ref, errs, status, info = oracle_decode(data, 0, chunk)
        except RuntimeError:
            stats["open_err"] += 1
            try:
                emul_decode_file(data, 0, chunk, 0)
                print("MISMATCH open: oracle refused, index accepted; iter", it)
                stats["mismatch"] += 1
            except RuntimeError:
                pass
            continue
        if status != 0:
            stats["oracle_exc"] += 1
            continue
        try:
            out, finfo, res, descs = emul_decode_file(data, 0, chunk, 0)
        except RuntimeError as e:
            print("MISMATCH open: index refused (%s), oracle decoded %d values; iter %d" % (e, ref.size, it))
            stats["mismatch"] += 1
            continue
        inexact = any(r.rflags & 8 for r in res) or finfo.stopped_early
        same = out.size == ref.size and np.array_equal(out, ref) and sum(1 for r in res if r.rflags & 1) == errs
        if same:
            stats["ok"] += 1
        elif inexact:
            stats["inexact"] += 1
        else:
            stats["mismatch"] += 1
            nd = int((out[:min(out.size, ref.size)] != ref[:min(out.size, ref.size)]).sum())
            print("MISMATCH iter %d chunk %d sizes %d/%d ndiff %d crc %d/%d flags %s" % (
                it, chunk, out.size, ref.size, nd, sum(1 for r in res if r.rflags & 1), errs, [hex(r.rflags) for r in res if r.rflags][:4]))
            with open("/tmp/fuzz_fail_%d.wv" % it, "wb") as f:
                f.write(data)
    print(stats, flush=True)
    return 1 if stats["mismatch"] else 0


if __name__ == "__main__":
    sys.exit(main())
