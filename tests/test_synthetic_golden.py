"""Committed hashes for the in-repo synthetic streams (tests/golden/synthetic_manifest.json, minted by
tools/make_golden_synthetic.py): one stream per BASELINE.json config, regenerated from its seed, and the stored DSD
fixtures that FFmpeg's independent DSD decoders vouched for.  The encoder, the oracle, the host-compiled device code and
(`-m gpu`) the CUDA path must all reproduce the minted values: PCM MD5, int32 MD5, per-block CRC list, error counts, getters."""
import hashlib
import json
import os

import numpy as np
import pytest

from _harness import emul_decode_file, format_samples, make_file, oracle_decode
from synthetic_configs import CONFIGS, DSD_FIXTURES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLD, "synthetic_manifest.json")) as f:
    MANIFEST = json.load(f)


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


_cache = {}


def stream(name):
    if name not in _cache:
        if name in CONFIGS:
            c = CONFIGS[name]
            _cache[name] = (make_file(seed=c["seed"], seconds=c["seconds"], **c["kw"])[2], c.get("open_flags", 0), MANIFEST["configs"][name])
        else:
            _cache[name] = (open(os.path.join(GOLD, name + ".wv"), "rb").read(), 0, MANIFEST["dsd_fixtures"][name])
    return _cache[name]


NAMES = sorted(CONFIGS) + sorted(DSD_FIXTURES)


def test_manifest_covers_every_baseline_config():
    assert sorted(MANIFEST["configs"]) == sorted(CONFIGS) and sorted(MANIFEST["dsd_fixtures"]) == sorted(DSD_FIXTURES)
    assert {e["baseline_config"] for e in MANIFEST["configs"].values()} == {1, 2, 3, 4, 5}
    e = MANIFEST["configs"]["config1_60s_s16_stereo"]
    assert e["samples"] == 2646000 and e["nblocks"] == 120
    for e in MANIFEST["dsd_fixtures"].values():  # what FFmpeg established at mint time
        assert e["ff_float_equals_raw_mode"] and e["raw_payload_is_source"] and e["ff_output_sensitive_to_one_bit"]


@pytest.mark.parametrize("name", NAMES)
def test_encoder_and_oracle_reproduce_the_minted_values(name):
    data, flags, e = stream(name)
    assert md5(data) == e["wv_md5"], "the synthetic encoder's output changed"
    out, errs, status, info = oracle_decode(data, flags, 4096)
    assert status == 0 and errs == e["crc_errors"] == 0
    assert md5(np.ascontiguousarray(out, dtype="<i4").tobytes()) == e["int32_md5"]
    assert md5(format_samples(out, info["bytes_per_sample"])) == e["pcm_md5"]
    for k, v in e["getters"].items():
        assert info[k] == v, k
    if "source_bytes_md5" in e:  # DSD fixture: the decode is the source byte stream FFmpeg's decoder also arrived at
        assert md5(out.astype(np.uint8).tobytes()) == e["source_bytes_md5"]


@pytest.mark.parametrize("name", [n for n in NAMES if n != "config1_60s_s16_stereo"])
def test_device_code_on_host_reproduces_the_minted_values(name):
    data, flags, e = stream(name)
    pcm, finfo, res, descs = emul_decode_file(data, flags, 4096, 1)
    assert md5(pcm) == e["pcm_md5"]
    assert not any(r.rflags for r in res)
    crcs = np.array([r.crc for r in res], dtype="<i4").view("<u4")
    if not flags:  # (OPEN_2CH_MAX decodes only the INITIAL block of every segment)
        assert md5(crcs.tobytes()) == e["block_crcs_md5"]


@pytest.mark.gpu
def test_cuda_path_reproduces_the_minted_values():
    """All streams in one batch per open-flag group through the C ABI: PCM MD5, per-block CRCs, zero flagged blocks."""
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import decode_files
    if N.load().wvb_device_count() <= 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    for flags in (0, 8):
        names = [n for n in NAMES if stream(n)[1] == flags]
        res = decode_files([stream(n)[0] for n in names], open_flags=flags, out_format=N.OUT_PCM)
        for n, (pcm, errs, info, results) in zip(names, res):
            e = stream(n)[2]
            assert errs == 0 and not any(r.rflags for r in results), n
            assert md5(pcm) == e["pcm_md5"], n
            if not flags:
                assert md5(np.array([r.crc for r in results], dtype="<i4").tobytes()) == e["block_crcs_md5"], n
        res = decode_files([stream(n)[0] for n in names], open_flags=flags, out_format=N.OUT_INT32)
        for n, (out, errs, info, results) in zip(names, res):
            assert md5(np.ascontiguousarray(out, dtype="<i4").tobytes()) == stream(n)[2]["int32_md5"], n


@pytest.mark.gpu
def test_config1_wvdemo_loop_on_device():
    """BASELINE configs[0]: the 60 s file decoded in 4096-sample calls like WvDemo.cs:110-135, PCM MD5 against the minted value."""
    from wavpackdecoder_b200 import wavpack_utils as W
    data, flags, e = stream("config1_60s_s16_stereo")
    wpc = W.WavpackOpenFileInput(data)
    buf = np.zeros(4096 * 2, dtype=np.int32)
    pcm = np.zeros(4096 * 4, dtype=np.uint8)
    h = hashlib.md5()
    total = 0
    while True:
        n = W.WavpackUnpackSamples(wpc, buf, 4096)
        if n == 0:
            break
        assert W.WavpackFormatSamples(buf, n * 2, 2, pcm)
        h.update(pcm[: n * 4].tobytes())
        total += n
    assert total == e["samples"] == W.WavpackGetNumSamples(wpc)
    assert h.hexdigest() == e["pcm_md5"] and W.WavpackGetNumErrors(wpc) == 0 and wpc.decode_passes == 1
