"""WavpackOpenFileInput + getters of the Python mirror (wavpackdecoder_b200.wavpack_utils) against the oracle's
WavpackContext for the same bytes.  Open and getters are host-only (index pass), so this runs without a GPU."""
import numpy as np
import pytest

from _harness import KIND_DSD, OracleFile, make_file
from cases import DSD_CASES, PCM_CASES
from wavpackdecoder_b200 import wavpack_utils as W

CASES = PCM_CASES + DSD_CASES


@pytest.mark.parametrize("name,flags,chunk,kw", CASES, ids=[c[0] for c in CASES])
def test_getters_match_oracle(name, flags, chunk, kw):
    kw = dict(kw)
    if "nsamples" not in kw:
        kw["seconds"] = min(kw.get("seconds", 1.0), 0.3)
    cfg, src, data = make_file(**kw)
    o = OracleFile(data, flags)
    assert o.error is None
    i = o.info()
    wpc = W.WavpackOpenFileInput(data, flags)
    assert W.WavpackGetErrorMessage(wpc) is None
    assert W.WavpackGetNumSamples(wpc) == i["num_samples"]
    assert W.WavpackGetNumSamples(wpc, True) == i["num_samples_native"]
    assert W.WavpackGetSampleRate(wpc) == i["sample_rate"]
    assert W.WavpackGetNumChannels(wpc) == i["num_channels"]
    assert W.WavpackGetReducedChannels(wpc) == i["reduced_channels"]
    assert W.WavpackGetBitsPerSample(wpc) == i["bits_per_sample"]
    assert W.WavpackGetBytesPerSample(wpc) == i["bytes_per_sample"]
    assert W.WavpackGetFileFormat(wpc) == i["file_format"]
    assert W.WavpackGetFileExtension(wpc) == i["file_extension"]
    assert W.WavpackGetIsFive(wpc) == i["is_five"]
    assert W.WavpackGetVersion(wpc) == i["version"]
    assert W.WavpackGetIsFloat(wpc) == i["is_float"]
    assert W.WavpackGetHeader(wpc) == i["header"]
    assert W.WavpackGetMode(wpc) == i["mode"]
    assert W.WavpackGetCompressionLevel(wpc) == i["compression_level"]
    assert W.WavpackLossy(wpc) == i["lossy"]
    assert W.WavpackGetSampleIndex(wpc) == 0 and W.WavpackGetNumErrors(wpc) == 0
    o.close()


def test_open_errors_match_oracle():
    cfg, src, six = make_file(channels=6, bits=16, seconds=0.2)
    for data, flags in [(b"", 0), (b"not a wavpack file" * 100, 0), (six, 0), (six[:40], 0)]:
        o = OracleFile(data, flags)
        wpc = W.WavpackOpenFileInput(data, flags)
        assert W.WavpackGetErrorMessage(wpc) == o.error, (o.error, W.WavpackGetErrorMessage(wpc))
        o.close()
    # bad metadata id in the first block (ID_ENCODER_INFO = 1 is not optional, MetadataUtils.cs:187-191)
    cfg, src, ok = make_file(seconds=0.1)
    bad = bytearray(ok)
    bad[32] = 0x01
    o = OracleFile(bytes(bad))
    wpc = W.WavpackOpenFileInput(bytes(bad))
    assert o.error is not None and W.WavpackGetErrorMessage(wpc) == o.error
    o.close()


def test_format_samples_matches_oracle():
    from _harness import format_samples
    rng = np.random.default_rng(7)
    v = rng.integers(-2 ** 31, 2 ** 31 - 1, size=1000, dtype=np.int64).astype(np.int32)
    for bps in (1, 2, 3, 4):
        for dsd in (False, True):
            pcm = np.zeros(v.size * bps, dtype=np.uint8)
            assert W.WavpackFormatSamples(v, v.size, bps, pcm, 0, dsd)
            assert np.array_equal(pcm, format_samples(v, bps, dsd and bps == 1))
    assert not W.WavpackFormatSamples(v, v.size, 2, np.zeros(10, dtype=np.uint8))
