#!/usr/bin/env python3
"""Mint tests/golden/*.wv with FFmpeg's native WavPack ENCODER (independent of the reference
and of this repo's encoder) and record what a correct decoder must produce.

Run once in the authoring container (needs tools/ffwv.py's bundled libavcodec):
    python tools/make_golden.py
Outputs: tests/golden/ff_*.wv and tests/golden/manifest.json.  For integer inputs the expected
PCM is the SOURCE fed to FFmpeg (ground truth independent of our oracle).  For float inputs the
reference only yields 24-bit mantissa integers, so the manifest records the
oracle's output hash as a regression value and relies on the block CRC (computed by FFmpeg from
its own integers) for independence.
"""
import hashlib
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
sys.path.insert(0, HERE)
import ffwv  # noqa: E402
from _harness import make_config, synth, oracle_decode, KIND_FLOAT  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def patch_total(wv, total):
    b = bytearray(wv)
    for off, size, h in ffwv.split_blocks(wv):
        struct.pack_into("<I", b, off + 12, total)
    return bytes(b)


def src_pcm(bits, nch, rate, seconds, seed):
    cfg = make_config(bits=bits, channels=nch, sample_rate=rate)
    return synth(cfg, seed, int(rate * seconds))


CASES = [
    # name, fmt, bits, nch, rate, seconds, compression_level, extra opts
    ("ff_s16_stereo_c0", "s16p", 16, 2, 44100, 0.7, 0, {}),
    ("ff_s16_stereo_c1", "s16p", 16, 2, 44100, 0.7, 1, {}),
    ("ff_s16_stereo_c2", "s16p", 16, 2, 44100, 0.7, 2, {}),
    ("ff_s16_stereo_c3", "s16p", 16, 2, 44100, 0.7, 3, {}),
    ("ff_s16_stereo_c5", "s16p", 16, 2, 44100, 0.6, 5, {}),
    ("ff_s16_stereo_c8", "s16p", 16, 2, 44100, 0.55, 8, {}),
    ("ff_s16_stereo_nojoint", "s16p", 16, 2, 44100, 0.6, 1, {"joint_stereo": "off"}),
    ("ff_s16_mono_c2", "s16p", 16, 1, 44100, 0.7, 2, {}),
    ("ff_u8_stereo_c1", "u8p", 8, 2, 22050, 0.8, 1, {}),
    ("ff_s24_stereo_c3", "s32p", 24, 2, 48000, 0.6, 3, {}),
    ("ff_s32_stereo_c1", "s32p", 32, 2, 44100, 0.6, 1, {}),
    ("ff_s32_mono_c4", "s32p", 32, 1, 44100, 0.6, 4, {}),
    ("ff_s16_optmono", "s16p", 16, 2, 44100, 0.6, 1, {"optimize_mono": "on"}),
    ("ff_flt_stereo_c1", "fltp", 32, 2, 44100, 0.6, 1, {}),
    ("ff_s16_51_c1", "s16p", 16, 6, 48000, 0.6, 1, {}),
    ("ff_s16_stereo_c1_long", "s16p", 16, 2, 44100, 3.2, 1, {}),  # seven blocks, incl. the generator's silence gap and L==R stretch
    ("ff_s24_mono_c1", "s32p", 24, 1, 48000, 0.9, 1, {}),
    ("ff_u8_mono_c0", "u8p", 8, 1, 22050, 1.1, 0, {}),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for i, (name, fmt, bits, nch, rate, secs, lvl, opts) in enumerate(CASES):
        src = src_pcm(min(bits, 24) if fmt == "fltp" else bits, nch, rate, secs, 0x60D0 + i)
        n = src.size // nch
        if name == "ff_s16_optmono":  # identical channels -> FALSE_STEREO blocks
            s2 = src.reshape(-1, 2)
            s2[:, 1] = s2[:, 0]
            src = s2.reshape(-1)
        feed = src
        if fmt == "fltp":
            feed = (src.astype(np.float64) / (1 << 23)).astype(np.float32)
        elif fmt == "s32p" and bits == 24:
            feed = src.astype(np.int32) << 8  # FFmpeg wants 24-bit left-justified in 32
        elif fmt == "u8p":
            feed = (src + 128).astype(np.uint8)
        o = dict(opts)
        if fmt == "s32p" and bits == 24:
            o["bits_per_raw_sample"] = 24
        wv = ffwv.ff_encode(feed, nch, rate, fmt, compression_level=lvl, opts=o)
        wv = patch_total(wv, n)
        flags = 0x8 if nch > 2 else 0
        out, errs, status, info = oracle_decode(wv, flags)
        rch = info["reduced_channels"]
        entry = dict(file=name + ".wv", fmt=fmt, src_bits=bits, channels=nch, rate=rate, samples=n, open_flags=flags,
                     reduced_channels=rch, bytes_per_sample=info["bytes_per_sample"], bits_per_sample=info["bits_per_sample"],
                     wv_md5=hashlib.md5(wv).hexdigest(),
                     blocks=[dict(samples=h["samples"], flags=h["flags"], crc=h["crc"]) for _, _, h in ffwv.split_blocks(wv)])
        if fmt == "fltp":
            entry["expected"] = "oracle-regression"
            exp = out
        else:
            entry["expected"] = "source"
            exp = src.reshape(-1, nch)[:, :rch].reshape(-1)
            if fmt == "s32p" and bits == 24 and info["bytes_per_sample"] == 4:
                exp = exp.astype(np.int32) << 8
        entry["int32_md5"] = hashlib.md5(np.ascontiguousarray(exp, dtype="<i4").tobytes()).hexdigest()
        same = out.size == exp.size and np.array_equal(out, exp)
        entry["oracle_ok_at_mint"] = bool(same and errs == 0 and status == 0)
        print("%-24s %7d B  blocks=%d  oracle_equal=%s crc_errors=%d status=%d bps=%d/%d lossy=%s" % (
            name, len(wv), len(entry["blocks"]), same, errs, status, info["bytes_per_sample"], info["bits_per_sample"], info["lossy"]))
        with open(os.path.join(OUT, name + ".wv"), "wb") as f:
            f.write(wv)
        manifest[name] = entry
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
