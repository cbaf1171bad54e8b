"""Integrity row: the WavPack 5 block checksum (ID_BLOCK_CHECKSUM, Defines.cs:83).  The reference only
notes the sub-block (MetadataUtils.cs:183); the batch decoder verifies it on the host (wvb_block_checksum_ok) and on the
device (k_block_checksum -> WVB_RF_BLOCK_CHECKSUM).  The definition is restated here in plain Python from WavPack 5's
published WavpackVerifySingleBlock, independently of the C and CUDA code under test."""
import struct

import numpy as np
import pytest

from _harness import X_BLOCK_CHECKSUM, X_CONFIG, X_RIFF_HEADER, format_samples, make_file, oracle_decode

KIND_DSD = 3


def blocks_of(data):
    """(offset, length) of every block of a well-formed stream."""
    at, out = 0, []
    while at + 32 <= len(data):
        assert data[at:at + 4] == b"wvpk"
        ln = struct.unpack_from("<I", data, at + 4)[0] + 8
        out.append((at, ln))
        at += ln
    return out


def py_block_checksum(block):
    """1 matching, 0 wrong, -1 none: walk the sub-blocks, sum the 16-bit words before the checksum sub-block."""
    at, end = 32, struct.unpack_from("<I", block, 4)[0] + 8
    while at + 2 <= end:
        ident, words, hdr = block[at], block[at + 1], 2
        if ident & 0x80:
            words |= (block[at + 2] << 8) | (block[at + 3] << 16)
            hdr = 4
        if (ident & 0x3f) == 0x2f:
            n = 2 * words
            if (ident & 0x40) or n not in (2, 4):
                return -1
            csum = 0xffffffff
            for i in range(0, at, 2):
                csum = (csum * 3 + block[i] + (block[i + 1] << 8)) & 0xffffffff
            stored = int.from_bytes(block[at + 2:at + 2 + n], "little")
            if n == 2:
                csum = (csum ^ (csum >> 16)) & 0xffff
            return int(stored == csum)
        at += hdr + 2 * words
    return -1


CASES = [
    ("s16", dict(seconds=0.6)),
    ("m24_odd_blocks", dict(bits=24, channels=1, block_samples=1001, seconds=0.3)),
    ("dsd_2byte", dict(kind=KIND_DSD, dsd_mode=1, seconds=0.05, block_samples=4000)),
    ("float", dict(kind=2, bits=32, seconds=0.2, block_samples=2000)),
]


def _host_ok(lib, block):
    buf = np.frombuffer(block, dtype=np.uint8)
    return lib.wvb_block_checksum_ok(buf.ctypes.data, buf.size)


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
def test_host_checksum_follows_the_published_definition(name, kw):
    from wavpackdecoder_b200 import _native as N
    lib = N.load()
    data = bytes(make_file(extras=X_RIFF_HEADER | X_CONFIG | X_BLOCK_CHECKSUM, **kw)[2])
    blks = blocks_of(data)
    assert len(blks) >= 2
    for off, ln in blks:
        blk = data[off:off + ln]
        assert struct.unpack_from("<I", blk, 24)[0] & 0x10000000  # HAS_CHECKSUM
        assert py_block_checksum(blk) == 1 and _host_ok(lib, blk) == 1
    # every single-byte change before the stored value is caught by both (the sum is injective in any one word)
    off, ln = blks[1]
    rng = np.random.default_rng(5)
    for pos in [8, 9, 31, 32, 40, ln - 7] + list(rng.integers(32, ln - 6, size=40)):
        blk = bytearray(data[off:off + ln])
        blk[int(pos)] ^= 1 << int(rng.integers(0, 8))
        assert py_block_checksum(bytes(blk)) == _host_ok(lib, bytes(blk)) != 1
    blk = bytearray(data[off:off + ln])
    blk[1] ^= 2  # not a block any more
    assert _host_ok(lib, bytes(blk)) == -1
    # files without the sub-block
    plain = bytes(make_file(**kw)[2])
    o, l = blocks_of(plain)[0]
    assert py_block_checksum(plain[o:o + l]) == -1 and _host_ok(lib, plain[o:o + l]) == -1
    assert _host_ok(lib, b"wvpk" + bytes(20)) == -1 and _host_ok(lib, b"") == -1


def test_index_marks_blocks_with_a_checksum():
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import Corpus
    with_ck = bytes(make_file(extras=X_CONFIG | X_BLOCK_CHECKSUM, seconds=0.5)[2])
    without = bytes(make_file(extras=X_CONFIG, seconds=0.5)[2])
    corpus = Corpus.from_files([with_ck, without])
    tab = N.desc_table(corpus.descs, corpus.nblocks)
    n0 = int(corpus.count[0])
    assert n0 == len(blocks_of(with_ck))
    assert all(tab["bflags"][:n0] & N.BF_BLOCK_CHECKSUM) and not any(tab["bflags"][n0:] & N.BF_BLOCK_CHECKSUM)
    for d, (off, ln) in zip(tab[:n0], blocks_of(with_ck)):
        at = int(d["in_offset"]) - int(corpus.offsets[0]) + int(d["checksum_off"])
        assert off < at < off + ln and with_ck[at] == 0x2f and with_ck[at + 1] == 2


@pytest.mark.gpu
def test_device_checksum_flags_exactly_the_damaged_blocks():
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import decode_files
    files, expect = [], []
    rng = np.random.default_rng(11)
    for name, kw in CASES:
        good = bytes(make_file(extras=X_RIFF_HEADER | X_CONFIG | X_BLOCK_CHECKSUM, **kw)[2])
        blks = blocks_of(good)
        files.append(good); expect.append([1] * len(blks))
        # damage block 1 in a place the decoder does not look at (the stored checksum itself), and block 0 in its audio
        bad = bytearray(good)
        o1, l1 = blks[1]
        bad[o1 + l1 - 1] ^= 0x40
        o0, l0 = blks[0]
        bad[o0 + l0 - 40] ^= 0x04
        files.append(bytes(bad)); expect.append([0, 0] + [1] * (len(blks) - 2))
    files.append(bytes(make_file(extras=0, seconds=0.2)[2])); expect.append(None)  # no checksums at all
    for fmt in (N.OUT_INT32, N.OUT_PCM):
        res = decode_files(files, out_format=fmt)
        for data, exp, (out, errs, info, results) in zip(files, expect, res):
            got = [0 if r.rflags & N.RF_BLOCK_CHECKSUM else 1 for r in results]
            if exp is None:
                assert all(got)
                continue
            host = [py_block_checksum(data[o:o + l]) for o, l in blocks_of(data)]
            assert got == exp == host
            # the check is an extension: samples and crc_errors stay the reference's
            ref, rerrs, status, rinfo = oracle_decode(data)
            if any(r.rflags & N.RF_INEXACT for r in results):
                continue
            if fmt == N.OUT_INT32:
                assert np.array_equal(out, ref) and errs == rerrs
            elif not rinfo.get("is_dsd"):
                assert np.array_equal(out, format_samples(ref, rinfo["bytes_per_sample"]))


@pytest.mark.gpu
def test_device_checksum_at_every_slab_alignment():
    """The kernel reads 16-byte vectors from the first aligned address; files packed at arbitrary byte offsets exercise
    the scalar head, the vector body, the partial last step and the odd-address fallback."""
    from wavpackdecoder_b200 import _native as N
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    base = bytes(make_file(extras=X_CONFIG | X_BLOCK_CHECKSUM, seconds=0.4, block_samples=3000)[2])
    small = bytes(make_file(extras=X_CONFIG | X_BLOCK_CHECKSUM, seconds=0.01, block_samples=50)[2])
    nb = len(blocks_of(base)) + len(blocks_of(small))
    l0 = blocks_of(base)[0][1]
    dec = BatchDecoder(0)
    try:
        for shift in range(0, 34):
            slab = np.zeros(shift + len(base) + len(small) + 128, dtype=np.uint8)
            slab[shift:shift + len(base)] = np.frombuffer(base, dtype=np.uint8)
            o2 = shift + len(base)
            slab[o2:o2 + len(small)] = np.frombuffer(small, dtype=np.uint8)
            damaged = shift % 3 == 1
            if damaged:
                slab[shift + l0 - 40] ^= 0x10  # inside block 0's audio bitstream
            corpus = Corpus(slab, [shift, o2], [len(base), len(small)], out_format=N.OUT_PCM)
            assert corpus.nblocks == nb
            _out, results = dec.decode_corpus(corpus)
            flagged = [i for i in range(nb) if results[i].rflags & N.RF_BLOCK_CHECKSUM]
            assert flagged == ([0] if damaged else []), (shift, flagged)
    finally:
        dec.close()


@pytest.mark.gpu
def test_verify_files_reports_block_checksum_errors():
    from wavpackdecoder_b200.batch import verify_files
    good = bytes(make_file(extras=X_RIFF_HEADER | X_CONFIG | X_BLOCK_CHECKSUM, seconds=0.6)[2])
    blks = blocks_of(good)
    bad = bytearray(good)
    bad[blks[1][0] + blks[1][1] - 1] ^= 0x40  # the stored checksum of block 1: the audio still decodes cleanly
    plain = bytes(make_file(extras=X_RIFF_HEADER | X_CONFIG, seconds=0.3)[2])
    res = verify_files([good, bytes(bad), plain])
    assert [r["block_checksum_errors"] for r in res] == [0, 1, None]
    assert [r["crc_errors"] for r in res] == [0, 0, 0]
