"""Container headers beyond the reference's RIFF: Wave64 and DSDIFF synthesis in
wavpackdecoder_b200.containers, parsed back here chunk by chunk from the published layouts, and the native-container mode
of wvdemo.unpack_files on the GPU.  tools/check_containers.py opens the same files with FFmpeg's w64 / iff demuxers."""
import struct

import numpy as np
import pytest

from _harness import KIND_DSD, make_file, oracle_decode, format_samples

X_RIFF, X_CONFIG, X_NEW_CONFIG = 1, 2, 4


def parse_w64(blob):
    """{guid-name: (offset of payload, payload length)} walking 8-byte aligned chunks; sizes include the 24-byte header."""
    from wavpackdecoder_b200 import containers as K
    assert blob[:16] == K.W64_RIFF and blob[24:40] == K.W64_WAVE
    total = struct.unpack_from("<Q", blob, 16)[0]
    assert total == len(blob)
    at, out = 40, {}
    while at < len(blob):
        guid, size = blob[at:at + 16], struct.unpack_from("<Q", blob, at + 16)[0]
        assert size >= 24
        out[guid[:4].decode()] = (at + 24, size - 24)
        at += (size + 7) & ~7
    assert at == len(blob)
    return out


def parse_iff(blob, at, end):
    out = []
    while at < end:
        cid, size = blob[at:at + 4], struct.unpack_from(">Q", blob, at + 4)[0]
        out.append((cid, at + 12, size))
        at += 12 + size + (size & 1)
    assert at == end
    return out


def test_w64_header_round_trip():
    from wavpackdecoder_b200 import containers as K
    for n, ch, rate, bits, byteps in [(1000, 2, 44100, 16, 2), (1001, 1, 48000, 24, 3), (7, 6, 96000, 24, 3), (0, 2, 44100, 16, 2)]:
        data = bytes(range(256)) * ((n * ch * byteps) // 256 + 1)
        data = data[:n * ch * byteps]
        blob = K.w64_header(n, ch, rate, bits, byteps) + data + K.w64_trailer(n, ch, byteps)
        assert len(K.w64_header(n, ch, rate, bits, byteps)) == 104
        chunks = parse_w64(blob)
        fo, fl = chunks["fmt "]
        assert fl == 16
        tag, c, r, brate, align, b = struct.unpack_from("<HHIIHH", blob, fo)
        assert (tag, c, r, brate, align, b) == (1, ch, rate, rate * ch * byteps, ch * byteps, bits)
        do, dl = chunks["data"]
        assert dl == len(data) and blob[do:do + dl] == data


def test_dff_header_round_trip():
    from wavpackdecoder_b200 import containers as K
    for n, ch, rate in [(4096, 2, 2822400), (4097, 1, 2822400), (333, 6, 5644800), (5, 3, 2822400)]:
        data = bytes((i * 7) & 0xff for i in range(n * ch))
        blob = K.dff_header(n, ch, rate) + data + K.dff_trailer(n, ch)
        assert blob[:4] == b"FRM8" and blob[12:16] == b"DSD "
        assert struct.unpack_from(">Q", blob, 4)[0] == len(blob) - 12
        top = {cid: (o, s) for cid, o, s in parse_iff(blob, 16, len(blob))}
        assert set(top) == {b"FVER", b"PROP", b"DSD "}
        assert struct.unpack_from(">I", blob, top[b"FVER"][0])[0] == 0x01050000
        po, ps = top[b"PROP"]
        assert blob[po:po + 4] == b"SND "
        prop = {cid: (o, s) for cid, o, s in parse_iff(blob, po + 4, po + ps)}
        assert struct.unpack_from(">I", blob, prop[b"FS  "][0])[0] == rate
        co, cs = prop[b"CHNL"]
        assert struct.unpack_from(">H", blob, co)[0] == ch and cs == 2 + 4 * ch
        mo, ms = prop[b"CMPR"]
        assert blob[mo:mo + 4] == b"DSD " and blob[mo + 4] == 14 and blob[mo + 5:mo + 19] == b"not compressed"
        do, ds = top[b"DSD "]
        assert ds == len(data) and blob[do:do + ds] == data


def test_caf_header_round_trip():
    from wavpackdecoder_b200 import containers as K
    for n, ch, rate, bits, byteps in [(1000, 2, 44100, 16, 2), (77, 1, 96000, 24, 3), (0, 6, 48000, 32, 4)]:
        h = K.caf_header(n, ch, rate, bits, byteps)
        assert len(h) == 68 and h[:4] == b"caff" and struct.unpack_from(">HH", h, 4) == (1, 0)
        assert h[8:12] == b"desc" and struct.unpack_from(">q", h, 12)[0] == 32
        sr, fid, flags, bpp, fpp, cpf, bpc = struct.unpack_from(">d4sIIIII", h, 20)
        assert (sr, fid, flags, bpp, fpp, cpf, bpc) == (float(rate), b"lpcm", 2, byteps * ch, 1, ch, bits)
        assert h[52:56] == b"data" and struct.unpack_from(">qI", h, 56) == (4 + n * ch * byteps, 0)


def dsf_layout(raw, frames, ch):
    """numpy restatement of the DSF data layout: per-channel 4096-byte blocks, bits reversed, last block zero padded."""
    a = np.frombuffer(raw, dtype=np.uint8).reshape(frames, ch)
    rev = np.array([int("{:08b}".format(i)[::-1], 2) for i in range(256)], dtype=np.uint8)
    groups = (frames + 4095) // 4096
    padded = np.zeros((groups * 4096, ch), dtype=np.uint8)
    padded[:frames] = rev[a]
    return padded.reshape(groups, 4096, ch).transpose(0, 2, 1).tobytes()


def test_dsf_header_layout():
    from wavpackdecoder_b200 import containers as K
    for n, ch, rate in [(4096, 2, 2822400), (4097, 1, 2822400), (10000, 6, 5644800), (1, 2, 2822400)]:
        h = K.dsf_header(n, ch, rate)
        data = K.dsf_data_bytes(n, ch)
        assert len(h) == 92 and data == (n + 4095) // 4096 * 4096 * ch
        assert struct.unpack_from("<4sQQQ", h, 0) == (b"DSD ", 28, 92 + data, 0)
        tag, size, ver, fid, ctype, nch, fs, bps, count, blk, rsv = struct.unpack_from("<4sQIIIIIIQII", h, 28)
        assert (tag, size, ver, fid, nch, fs, bps, count, blk, rsv) == (b"fmt ", 52, 1, 0, ch, rate, 1, n * 8, 4096, 0)
        assert ctype == {1: 1, 2: 2, 6: 7}[ch]
        assert struct.unpack_from("<4sQ", h, 80) == (b"data", 12 + data)


def _with_file_format(data, fmt):
    """Set the file_format byte of the ID_NEW_CONFIG_BLOCK the synthetic encoder writes (one payload byte, odd-size flag)."""
    data = bytearray(data)
    at = data.find(b"\x6a\x01")
    assert at > 0
    data[at + 2] = fmt
    return bytes(data)


@pytest.mark.gpu
def test_native_containers_on_device():
    from wavpackdecoder_b200 import containers as K
    from wavpackdecoder_b200.wvdemo import unpack_files
    pcm16 = _with_file_format(make_file(extras=X_CONFIG | X_NEW_CONFIG, seconds=0.7)[2], K.WP_FORMAT_W64)
    pcm24 = _with_file_format(make_file(extras=X_CONFIG | X_NEW_CONFIG, bits=24, channels=1, nsamples=10001)[2], K.WP_FORMAT_W64)
    wav = bytes(make_file(extras=X_CONFIG | X_NEW_CONFIG, seconds=0.3)[2])
    stored = bytes(make_file(extras=X_RIFF | X_CONFIG, seconds=0.3)[2])
    caf = _with_file_format(make_file(extras=X_CONFIG | X_NEW_CONFIG, bits=24, seconds=0.2)[2], K.WP_FORMAT_CAF)
    dsd = _with_file_format(make_file(kind=KIND_DSD, dsd_mode=1, extras=X_CONFIG | X_NEW_CONFIG, seconds=0.05, block_samples=4000)[2], K.WP_FORMAT_DFF)
    dsd_mono = _with_file_format(make_file(kind=KIND_DSD, dsd_mode=3, channels=1, extras=X_CONFIG | X_NEW_CONFIG, nsamples=4097, block_samples=2000)[2], K.WP_FORMAT_DFF)
    dsf = make_file(kind=KIND_DSD, dsd_mode=1, extras=X_CONFIG | X_NEW_CONFIG, nsamples=9000, block_samples=4000)[2]  # (the encoder writes DSF)
    dsf_mono = make_file(kind=KIND_DSD, dsd_mode=0, channels=1, extras=X_CONFIG | X_NEW_CONFIG, nsamples=4096, block_samples=2048)[2]
    files = [pcm16, pcm24, wav, stored, dsd, dsd_mono, caf, bytes(dsf), bytes(dsf_mono)]
    res = unpack_files(files, container="native")
    assert [code for _b, code in res] == [0] * len(files)
    for data, (blob, _code), kind in zip(files, res, ["w64", "w64", "wav", "stored", "dff", "dff", "caf", "dsf", "dsf"]):
        ref, errs, status, info = oracle_decode(data)
        assert status == 0 and errs == 0
        if kind == "w64":
            chunks = parse_w64(blob)
            do, dl = chunks["data"]
            assert blob[do:do + dl] == format_samples(ref, info["bytes_per_sample"]).tobytes()
            assert struct.unpack_from("<HHI", blob, chunks["fmt "][0]) == (1, info["reduced_channels"], info["sample_rate"])
        elif kind == "dsf":
            raw = np.asarray(ref, dtype=np.uint8).tobytes()
            ch = info["reduced_channels"]
            frames = len(raw) // ch
            assert blob[:92] == K.dsf_header(frames, ch, info["sample_rate"])
            assert blob[92:] == dsf_layout(raw, frames, ch)
        elif kind == "caf":
            assert blob[:4] == b"caff" and blob[68:] == format_samples(ref, info["bytes_per_sample"]).tobytes()
            assert struct.unpack_from(">q", blob, 56)[0] == 4 + len(blob) - 68
        elif kind in ("wav", "stored"):
            assert blob[:4] == b"RIFF" and blob[44:] == format_samples(ref, info["bytes_per_sample"]).tobytes()
        else:
            top = {cid: (o, s) for cid, o, s in parse_iff(blob, 16, len(blob))}
            do, ds = top[b"DSD "]
            assert blob[do:do + ds] == np.asarray(ref, dtype=np.uint8).tobytes()  # raw DSD bytes, not the demo's offset-binary
            po, ps = top[b"PROP"]
            prop = {cid: (o, s) for cid, o, s in parse_iff(blob, po + 4, po + ps)}
            assert struct.unpack_from(">I", blob, prop[b"FS  "][0])[0] == info["sample_rate"]
    with pytest.raises(NotImplementedError):  # PCM audio in a file that names a DSD container
        unpack_files([_with_file_format(pcm16, K.WP_FORMAT_DSF)], container="native")
