"""SetSample / SetTime (WavPackUtils.cs:502-594) through the API mirror against the oracle's restated seek():
return value, WavpackGetSampleIndex, the samples of the following WavpackUnpackSamples calls and WavpackGetNumErrors.
The CPU variant runs the mirror's window logic over the host-compiled device code (test seam); the `gpu` variant is the
product path."""
import ctypes as C

import numpy as np
import pytest

from _harness import KIND_DSD, KIND_HYBRID, EmulDecoder, OracleFile, make_file, refdec
from cases import corrupt_cases
from wavpackdecoder_b200 import wavpack_utils as W


def _lib():
    lib = refdec()
    lib.rd_set_sample.argtypes = [C.c_void_p, C.c_long]
    lib.rd_set_time.argtypes = [C.c_void_p, C.c_long]
    return lib


def _streams():
    out = [("stereo16", make_file(seconds=2.2)[2]),
           ("mono24", make_file(seconds=1.1, channels=1, bits=24, block_samples=7001)[2]),
           ("hybrid", make_file(seconds=1.0, kind=KIND_HYBRID)[2]),
           ("int32_wvx", make_file(seconds=0.8, bits=32, int32_sent_bits=8)[2]),
           ("dsd_fast", make_file(seconds=0.1, kind=KIND_DSD, dsd_mode=1, block_samples=8192)[2]),
           ("dsd_high", make_file(seconds=0.1, kind=KIND_DSD, dsd_mode=3, block_samples=8192)[2])]
    cc = {c[0]: c[1] for c in corrupt_cases()}
    out += [("damaged_flip_bitstream", cc["flip_bitstream"]), ("damaged_flip_header_crc", cc["flip_header_crc"]),
            ("damaged_dsd3_flip", cc["dsd3_flip"])]
    return out


STREAMS = _streams()


def _check_seek(data, make_decoder, targets, chunks=(4096, 1000), by_time=False):
    lib = _lib()
    for tgt in targets:
        for chunk in chunks:
            o = OracleFile(data)
            wpc = W.WavpackOpenFileInput(data)
            if make_decoder:
                wpc._dec = make_decoder()
            nch = W.WavpackGetReducedChannels(wpc)
            # a caller that has already read something, then seeks
            buf = np.zeros(chunk * nch, dtype=np.int32)
            n0 = W.WavpackUnpackSamples(wpc, buf, chunk)
            r0, rbuf = o.unpack(chunk, nch)
            assert n0 == r0 and np.array_equal(buf[: n0 * nch], rbuf[: r0 * nch])
            if by_time:
                ok, rok = W.SetTime(wpc, tgt), lib.rd_set_time(o.ctx, tgt)
            else:
                ok, rok = W.SetSample(wpc, tgt), lib.rd_set_sample(o.ctx, tgt)
            assert rok in (0, 1), (tgt, rok)
            assert bool(ok) == bool(rok), (tgt, ok, rok)
            if not rok:
                o.close()
                continue
            assert W.WavpackGetSampleIndex(wpc) == lib.rd_get_sample_index(o.ctx), tgt
            for _ in range(4):
                n = W.WavpackUnpackSamples(wpc, buf, chunk)
                rn, rbuf = o.unpack(chunk, nch)
                assert n == rn, (tgt, chunk, n, rn)
                assert np.array_equal(buf[: n * nch], rbuf[: n * nch]), (tgt, chunk)
                assert W.WavpackGetSampleIndex(wpc) == lib.rd_get_sample_index(o.ctx)
                assert W.WavpackGetNumErrors(wpc) == lib.rd_get_num_errors(o.ctx), (tgt, chunk)
            o.close()


def _targets(data):
    wpc = W.WavpackOpenFileInput(data)
    total = W.WavpackGetNumSamples(wpc)
    import struct
    starts, off = [], 0
    while off + 32 <= len(data) and data[off:off + 4] == b"wvpk":
        cks, = struct.unpack_from("<I", data, off + 4)
        bi, bs = struct.unpack_from("<II", data, off + 16)
        if bs:
            starts.append(bi)
        off += cks + 8
    t = {0, 1, -7, total - 1, total, total + 5, total // 2, total // 3 + 17}
    for s in starts[1:4]:
        t |= {s - 1, s, s + 1, s + 4095, s + 2048, s + 2049}
    return sorted(x for x in t if x < total + 10)


@pytest.mark.parametrize("name,data", STREAMS, ids=[s[0] for s in STREAMS])
def test_seek_matches_oracle_emulated(name, data):
    _check_seek(data, EmulDecoder, _targets(data))


def test_set_time_matches_oracle_emulated():
    data = STREAMS[0][1]
    _check_seek(data, EmulDecoder, [0, 999, 1000, 1500, 2000, 2199, 2200, 5000, -3], chunks=(4096,), by_time=True)


def test_seek_decodes_from_the_containing_block_only():
    """The window after a seek starts at the block that contains the target, not at the start of the file."""
    data = STREAMS[0][1]
    wpc = W.WavpackOpenFileInput(data)
    wpc._dec = EmulDecoder()
    assert W.SetSample(wpc, 3 * 22050 + 100)
    buf = np.zeros(8192, dtype=np.int32)
    assert W.WavpackUnpackSamples(wpc, buf, 4096) == 4096
    assert wpc._win.first == 3 * 22050 and wpc.decode_passes == 1
    # a bounded lookahead decodes that many blocks per pass
    wpc = W.WavpackOpenFileInput(data)
    wpc._dec = EmulDecoder()
    wpc.lookahead_blocks = 2
    got = []
    while True:
        n = W.WavpackUnpackSamples(wpc, buf, 4096)
        if n == 0:
            break
        got.append(buf[: n * 2].copy())
    o = OracleFile(data)
    ref, errs, status = o.decode_all(4096)
    o.close()
    assert np.array_equal(np.concatenate(got), ref) and wpc.decode_passes == 3  # 5 blocks, 2 per pass


def test_changing_the_call_size_rebuilds_the_grid():
    """A later WavpackUnpackSamples call with a different size decodes a new window with that call grid: on a damaged
    stream the mute starts where the reference's does."""
    cc = {c[0]: c[1] for c in corrupt_cases()}
    for data in (cc["flip_bitstream"], STREAMS[0][1]):
        o = OracleFile(data)
        wpc = W.WavpackOpenFileInput(data)
        wpc._dec = EmulDecoder()
        for chunk in (4096, 4096, 1000, 1000, 333, 4096, 50000, 4096):
            buf = np.zeros(chunk * 2, dtype=np.int32)
            n = W.WavpackUnpackSamples(wpc, buf, chunk)
            rn, rbuf = o.unpack(chunk, 2)
            assert n == rn
            assert np.array_equal(buf[: n * 2], rbuf[: n * 2]), chunk
            assert W.WavpackGetNumErrors(wpc) == o.lib.rd_get_num_errors(o.ctx)
        o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,data", STREAMS, ids=[s[0] for s in STREAMS])
def test_seek_matches_oracle_on_device(name, data):
    from wavpackdecoder_b200 import _native as N
    if N.load().wvb_device_count() <= 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    _check_seek(data, None, _targets(data))


@pytest.mark.gpu
def test_set_time_on_device():
    _check_seek(STREAMS[0][1], None, [0, 1000, 1500, 2199, 2200, 5000], chunks=(4096,), by_time=True)
