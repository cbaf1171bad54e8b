"""CPU-side regression of the DEVICE decode function: tests/emul compiles wavpackdecoder_b200/csrc/wvb_pcm.cuh
for the host and runs it block by block behind the real index pass (wvb_index).  This is a development
harness for boxes without a GPU (it is not shipped and not reachable from the product); the parity tests
proper are the `-m gpu` ones."""
import numpy as np
import pytest

from _harness import emul_decode_file, format_samples, make_file, oracle_decode
from cases import DSD_CASES, PCM_CASES, corrupt_cases


@pytest.mark.parametrize("name,flags,chunk,kw", PCM_CASES + DSD_CASES, ids=[c[0] for c in PCM_CASES + DSD_CASES])
def test_device_code_on_host_matches_oracle(name, flags, chunk, kw):
    kw = dict(kw)
    if "nsamples" not in kw:
        kw["seconds"] = min(kw.get("seconds", 1.0), 0.6)
    cfg, src, data = make_file(**kw)
    ref, errs, status, info = oracle_decode(data, flags, chunk)
    assert status == 0
    out, finfo, res, descs = emul_decode_file(data, flags, chunk, 0)
    assert out.size == ref.size
    if name in INEXACT_BY_DESIGN:  # state inherited from an earlier block: must be FLAGGED, need not match
        assert any(r.rflags & 8 for r in res)
    else:
        assert sum(1 for r in res if r.rflags & 1) == errs
        assert np.array_equal(out, ref)
    assert not any(r.rflags & (8 | 16) for r in res)
    pcm, _, _, _ = emul_decode_file(data, flags, chunk, 1)
    assert np.array_equal(pcm, format_samples(ref, info["bytes_per_sample"]))
    # getters mirrored by the index pass
    assert finfo.total_samples == info["num_samples"]
    assert finfo.num_channels == info["num_channels"]
    assert finfo.bytes_per_sample == info["bytes_per_sample"]
    assert (finfo.bits_per_sample // 8 if finfo.dsd_multiplier else finfo.bits_per_sample) == info["bits_per_sample"]  # WavPackUtils.cs:412
    assert bool(finfo.five) == info["is_five"]
    assert finfo.version == info["version"]


CORRUPT = corrupt_cases()
from cases import INEXACT_BY_DESIGN  # noqa: E402


@pytest.mark.parametrize("name,data,flags,chunk", CORRUPT, ids=[c[0] for c in CORRUPT])
def test_damaged_streams_follow_the_oracle(name, data, flags, chunk):
    ref, errs, status, info = oracle_decode(data, flags, chunk)
    assert status == 0
    out, finfo, res, descs = emul_decode_file(data, flags, chunk, 0)
    assert out.size == ref.size
    if name in INEXACT_BY_DESIGN:  # state inherited from an earlier block: must be FLAGGED, need not match
        assert any(r.rflags & 8 for r in res)
    else:
        assert sum(1 for r in res if r.rflags & 1) == errs
        assert np.array_equal(out, ref)

