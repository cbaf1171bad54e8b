#!/usr/bin/env python3
"""Mint tests/golden/synthetic_manifest.json (+ the small dsd_*.wv fixtures): committed hashes for streams of the in-repo
synthetic encoder, one per BASELINE.json config, so that a simultaneous regression of encoder and oracle cannot pass
unnoticed (SURVEY.md 8c last row), and an INDEPENDENT pin of the DSD decoders.

Run once in the authoring container (needs tools/ffwv.py's bundled libavcodec):
    python tools/make_golden_synthetic.py

Per entry the manifest holds what a correct decoder must produce (PCM MD5 as WavpackFormatSamples packs it, int32 MD5,
per-block CRC list, crc_errors, lossy, getters) and how the value was pinned at mint time:
  * lossless integer PCM: the decode equals the SOURCE signal, and FFmpeg's native decoder decodes the same stream to
    the same samples (`ff_decode_equal`);
  * hybrid: FFmpeg's decoder agrees with the oracle on the lossy reconstruction where it implements the mode (mono and
    HYBRID_BALANCE stereo; it ignores nothing else);
  * float-flagged: oracle regression value + source CRCs (the reference returns shifted/clipped 24-bit integers, FFmpeg
    returns floats);
  * DSD modes 0/1/3: FFmpeg's decoder has its own implementations of the "fast" and "high" DSD decoders.  It hands out
    PCM floats (its dsd2pcm filter over the decoded DSD bytes), so the pin is: FFmpeg's float output for the mode-1 and
    mode-3 streams is bit-identical to its output for the RAW (mode 0) stream of the same source, whose payload is the
    source bytes verbatim, while one flipped payload bit changes that output.  The oracle must return exactly those
    source bytes.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
sys.path.insert(0, HERE)
import ffwv  # noqa: E402
from _harness import KIND_DSD, KIND_FLOAT, KIND_HYBRID, OracleFile, format_samples, make_file, oracle_decode  # noqa: E402
from synthetic_configs import CONFIGS, DSD_FIXTURES  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


def describe(data, flags, chunk=4096):
    """What the oracle returns for a stream: hashes, per-block CRCs, getters."""
    out, errs, status, info = oracle_decode(data, flags, chunk)
    bps = info["bytes_per_sample"]
    dsd = bool(info["mode"] & 0x10000)
    pcm = format_samples(out, bps, dsd=False)
    blocks = [dict(samples=h["samples"], crc=h["crc"]) for _, _, h in ffwv.split_blocks(data) if h["samples"]]
    getters = {k: info[k] for k in ("num_samples", "num_samples_native", "sample_rate", "num_channels", "reduced_channels", "bits_per_sample",
                                    "bytes_per_sample", "lossy", "file_format", "file_extension", "is_five", "version", "is_float", "mode",
                                    "compression_level")}
    return out, dict(wv_md5=md5(data), wv_bytes=len(data), samples=int(out.size // max(info["reduced_channels"], 1)), status=status, crc_errors=errs,
                     int32_md5=md5(np.ascontiguousarray(out, dtype="<i4").tobytes()), pcm_md5=md5(pcm), nblocks=len(blocks),
                     block_crcs_md5=md5(np.array([b["crc"] for b in blocks], dtype="<u4").tobytes()),
                     first_block_crcs=[b["crc"] for b in blocks[:8]], getters=getters), dsd


def main():
    manifest = {"configs": {}, "dsd_fixtures": {}}
    for name, c in CONFIGS.items():
        cfg, src, data = make_file(seed=c["seed"], seconds=c["seconds"], **c["kw"])
        out, entry, dsd = describe(data, c.get("open_flags", 0))
        entry.update(baseline_config=c["baseline_config"], seed=c["seed"], seconds=c["seconds"], kw=c["kw"], open_flags=c.get("open_flags", 0))
        nch = cfg.channels
        rch = entry["getters"]["reduced_channels"]
        pins = []
        if cfg.kind == KIND_DSD:
            want = src.reshape(-1, nch)[:, :rch].reshape(-1)
            assert np.array_equal(out, want), name
            pins.append("decode == source DSD bytes")
        elif cfg.kind == 0 and not (cfg.bits == 32 and cfg.int32_sent_bits and not cfg.int32_wvx):
            want = src.reshape(-1, nch)[:, :rch].reshape(-1)
            assert np.array_equal(out, want), name
            pins.append("decode == source PCM")
        if cfg.kind in (0, KIND_HYBRID) and nch <= 2 and not (cfg.kind == KIND_HYBRID and nch == 2 and not cfg.hybrid_balance):
            ff, fmt = ffwv.ff_decode(data, nch)
            ffv = ff.reshape(-1).astype(np.int64)
            if fmt == "s32p":
                ffv = ffv >> (32 - 8 * entry["getters"]["bytes_per_sample"]) if entry["getters"]["bytes_per_sample"] < 4 else ffv
            if fmt == "u8p":
                ffv = ffv - 128
            same = ffv.size == out.size and np.array_equal(ffv.astype(np.int32), out)
            entry["ff_decode_equal"] = bool(same)
            assert same, (name, fmt)
            pins.append("FFmpeg libavcodec 62.11 native decoder output identical")
        entry["pinned_by"] = pins or ["oracle regression value; block CRCs computed by the encoder from the source"]
        assert entry["status"] == 0 and entry["crc_errors"] == 0, name
        manifest["configs"][name] = entry
        print("%-28s %9d B %4d blocks  %s" % (name, len(data), entry["nblocks"], "; ".join(entry["pinned_by"])))

    # ---- DSD: committed streams + the FFmpeg cross-check ----
    for name, c in DSD_FIXTURES.items():
        kw = dict(c["kw"])
        cfg, src, data = make_file(seed=c["seed"], seconds=c["seconds"], **kw)
        kw0 = dict(kw, dsd_mode=0)
        _, src0, raw = make_file(seed=c["seed"], seconds=c["seconds"], **kw0)
        assert np.array_equal(src, src0)
        nch = cfg.channels
        ff, fmt = ffwv.ff_decode(data, nch)
        ff0, _ = ffwv.ff_decode(raw, nch)
        assert fmt == "fltp" and ff.shape == ff0.shape and ff.shape[0] == src.size // nch
        equal = ff.tobytes() == ff0.tobytes()
        # the raw stream's payloads are the source bytes verbatim
        pay = bytearray()
        for off, size, h in ffwv.split_blocks(raw):
            blk = raw[off:off + size]
            at = 32
            while at < len(blk):
                idb, words, hdr = blk[at], blk[at + 1], 2
                if idb & 0x80:
                    words |= (blk[at + 2] << 8) | (blk[at + 3] << 16)
                    hdr = 4
                ln = words * 2 - (1 if idb & 0x40 else 0)
                if (idb & 0x3f) == 0x0e:
                    pay += blk[at + hdr + 2:at + hdr + ln]
                at += hdr + words * 2
        verbatim = bytes(pay) == src.astype(np.uint8).tobytes()
        # sensitivity: one flipped DSD bit in the raw payload changes FFmpeg's output
        first = next(ffwv.split_blocks(raw))
        mid = first[0] + first[1] // 2
        bad = bytearray(raw)
        bad[mid] ^= 0x10
        try:
            ffb, _ = ffwv.ff_decode(bytes(bad), nch)
            sensitive = ffb.tobytes() != ff0.tobytes()
        except Exception:
            sensitive = True
        out, entry, _ = describe(data, 0)
        assert np.array_equal(out, src) and entry["crc_errors"] == 0
        assert equal and verbatim and sensitive, (name, equal, verbatim, sensitive)
        entry.update(file=name + ".wv", seed=c["seed"], seconds=c["seconds"], kw=kw, channels=nch,
                     source_bytes_md5=md5(src.astype(np.uint8).tobytes()), ff_float_md5=md5(ff.tobytes()),
                     ff_float_equals_raw_mode=bool(equal), raw_payload_is_source=bool(verbatim), ff_output_sensitive_to_one_bit=bool(sensitive),
                     pinned_by=["FFmpeg libavcodec 62.11 DSD decoder: PCM image identical to that of the raw-mode stream of the same source",
                                "decode == source DSD bytes"])
        with open(os.path.join(OUT, name + ".wv"), "wb") as f:
            f.write(data)
        manifest["dsd_fixtures"][name] = entry
        print("%-28s %9d B %4d blocks  ff==raw:%s verbatim:%s sensitive:%s" % (name, len(data), entry["nblocks"], equal, verbatim, sensitive))

    with open(os.path.join(OUT, "synthetic_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
