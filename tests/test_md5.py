"""Integrity row: the stored ID_MD5_CHECKSUM lookup (host) and MD5 of decoded PCM on the device
(wvb_batch_md5) against hashlib."""
import hashlib

import numpy as np
import pytest

from _harness import make_file, oracle_decode, format_samples

X_RIFF, X_CONFIG, X_MD5_TRAILER = 1, 2, 16
FAKE_MD5 = bytes((i * 17 + 3) & 0xff for i in range(16))  # what the synthetic encoder writes (corpus/wvenc.c)


def _with_real_md5(kw):
    """A synthetic file whose ID_MD5_CHECKSUM holds the MD5 of its source PCM (the encoder writes a placeholder)."""
    cfg, src, data = make_file(extras=X_RIFF | X_CONFIG | X_MD5_TRAILER, **kw)
    data = bytes(data)
    samples, _errs, _status, info = oracle_decode(data, 0, 4096)
    pcm = format_samples(samples, info["bytes_per_sample"]).tobytes()
    at = data.find(FAKE_MD5)
    assert at > 0 and data.find(FAKE_MD5, at + 1) < 0
    digest = hashlib.md5(pcm).digest()
    return data[:at] + digest + data[at + 16:], pcm, digest


def test_stored_md5_lookup():
    from wavpackdecoder_b200.batch import stored_md5
    data, _pcm, digest = _with_real_md5(dict())
    assert stored_md5(data) == digest
    assert stored_md5(bytes(make_file(extras=X_RIFF)[2])) is None
    assert stored_md5(data[:len(data) - 40]) is None  # the final block (which carries it) truncated away
    assert stored_md5(b"") is None and stored_md5(b"wvpk" + bytes(40)) is None


@pytest.mark.gpu
def test_device_md5_matches_hashlib_on_arbitrary_ranges():
    import torch
    from wavpackdecoder_b200.batch import BatchDecoder
    rng = np.random.default_rng(7)
    blob = rng.integers(0, 256, size=1 << 20, dtype=np.uint8)
    d = torch.from_numpy(blob).cuda()
    offs, lens = [], []
    for ln in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 121, 127, 128, 1000, 4096, 65537, 300001]:
        for o in [0, 1, 2, 3, 64, 4099]:
            offs.append(o); lens.append(ln)
    dec = BatchDecoder(0)
    try:
        dig = dec.md5_ranges(offs, lens, blob.size, d.data_ptr())
    finally:
        dec.close()
    for o, ln, g in zip(offs, lens, dig):
        assert g.tobytes() == hashlib.md5(blob[o:o + ln].tobytes()).digest(), (o, ln)


@pytest.mark.gpu
def test_verify_files_against_stored_md5():
    from wavpackdecoder_b200.batch import verify_files
    good16, pcm16, dig16 = _with_real_md5(dict())
    good24, _p, dig24 = _with_real_md5(dict(bits=24, channels=1, block_samples=1001, seconds=0.3))
    good8, _p, dig8 = _with_real_md5(dict(bits=8))
    nomd5 = bytes(make_file(extras=X_RIFF)[2])
    bad = bytearray(good16)
    bad[len(bad) // 2] ^= 0x21  # damaged audio: CRC error in one block, MD5 of the (muted) output no longer matches
    res = verify_files([good16, good24, good8, nomd5, bytes(bad)])
    assert [r["match"] for r in res] == [True, True, True, None, False]
    assert res[0]["md5"] == dig16.hex() == hashlib.md5(pcm16).hexdigest()
    assert res[1]["md5"] == dig24.hex() and res[2]["md5"] == dig8.hex()
    assert [r["crc_errors"] > 0 for r in res] == [False, False, False, False, True]
