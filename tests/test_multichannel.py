"""SURVEY 8f-2 (extension): WVB_OPEN_ALL_CHANNELS indexes every block of a multichannel segment and the decoder
interleaves all channels.  The reference API only reaches the first stereo pair (OPEN_2CH_MAX); the other blocks are
pinned by (a) the source PCM (lossless, CRC from source) and (b) the oracle's per-block routine run directly on each
block (rd_dbg_unpack_current_block)."""
import ctypes as C
import struct

import numpy as np
import pytest

from _harness import OracleFile, emul_decode_file, format_samples, make_file
from cases import T16
from wavpackdecoder_b200 import _native as N

KW = dict(channels=6, bits=24, sample_rate=48000, block_samples=24000, terms=T16, deltas=[2] * 16, seconds=1.2)


def _blocks(data):
    off, out = 0, []
    while off + 32 <= len(data):
        cks, = struct.unpack_from("<I", data, off + 4)
        nsamp, flags = struct.unpack_from("<II", data, off + 20)
        out.append((off, cks + 8, nsamp, flags))
        off += cks + 8
    return out


def _oracle_block(data, off, nsamp, mono):
    f = OracleFile(data[off:], 8)
    assert f.error is None
    nch = 1 if mono else 2
    buf = np.zeros(nsamp * nch, dtype=np.int32)
    n = f.lib.rd_dbg_unpack_current_block(f.ctx, buf.ctypes.data, buf.size, 4096)
    assert n == nsamp and not f.lib.rd_dbg_check_crc_error(f.ctx)
    f.close()
    return buf.reshape(-1, nch)


def test_all_channels_emulated_device_code():
    cfg, src, data = make_file(**KW)
    out, info, res, descs = emul_decode_file(data, N.OPEN_ALL_CHANNELS, 4096, 0)
    assert info.num_channels == 6 and len(res) == 12 and not any(r.rflags for r in res)
    assert np.array_equal(out, src)
    got = out.reshape(-1, 6)
    ch = 0
    for off, size, nsamp, flags in _blocks(data):
        if flags & 0x800:
            ch = 0
        mono = bool(flags & 4)
        index, = struct.unpack_from("<I", data, off + 16)
        blk = _oracle_block(data, off, nsamp, mono)
        assert np.array_equal(got[index:index + nsamp, ch:ch + blk.shape[1]], blk)
        ch += blk.shape[1]
    pcm, _, _, _ = emul_decode_file(data, N.OPEN_ALL_CHANNELS, 4096, 1)
    assert np.array_equal(pcm, format_samples(src, 3))


@pytest.mark.gpu
def test_all_channels_gpu():
    from wavpackdecoder_b200.batch import decode_files
    cfg, src, data = make_file(**KW)
    cfg2, src2, data2 = make_file(seed=77, **dict(KW, bits=16, terms=[18, 18, 2, 3, -2], deltas=[2] * 5))
    res = decode_files([data, data2], open_flags=N.OPEN_ALL_CHANNELS, out_format=N.OUT_INT32)
    for (out, errs, info, results), s in zip(res, (src, src2)):
        assert errs == 0 and info.num_channels == 6
        assert np.array_equal(out, s)
