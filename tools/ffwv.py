#!/usr/bin/env python3
"""ctypes access to the WavPack encoder/decoder inside the libavcodec that ships in
the opencv-python-headless wheel of THIS container (FFmpeg libavcodec 62.11.100).

Used only here, offline, to (a) mint golden .wv fixtures from an implementation that is
independent of both the reference and this repo, and (b) check that the in-repo synthetic
encoder emits streams an independent decoder accepts.  Nothing in tests/, bench.py or the
product imports this at run time on the GPU box (the wheel's private .so layout is not a
stable interface); the fixtures it makes are committed under tests/golden/.
"""
import ctypes as C
import glob
import os
import struct

import numpy as np

_LIBDIR = None
for cand in glob.glob("/opt/prime-rl/.venv/lib/python3*/site-packages/opencv_python_headless.libs"):
    _LIBDIR = cand

_libs = {}


def _load():
    if _libs:
        return _libs
    if _LIBDIR is None:
        raise RuntimeError("no bundled libavcodec found")
    # dependency order; RTLD_GLOBAL so later libs resolve against earlier ones
    order = ["libcrypto", "libssl", "libdrm", "libpng16", "libaom", "libvpx", "libavutil", "libswresample", "libavcodec"]
    for n in order:
        hits = glob.glob(os.path.join(_LIBDIR, n + "-*"))
        if hits:
            try:
                _libs[n] = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
            except OSError:
                pass
    av, u = _libs["libavcodec"], _libs["libavutil"]
    for f in ("avcodec_find_encoder_by_name", "avcodec_find_decoder_by_name", "avcodec_alloc_context3", "av_packet_alloc", ):
        getattr(av, f).restype = C.c_void_p
    av.avcodec_find_encoder_by_name.argtypes = [C.c_char_p]
    av.avcodec_find_decoder_by_name.argtypes = [C.c_char_p]
    av.avcodec_alloc_context3.argtypes = [C.c_void_p]
    av.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    av.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    av.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    av.avcodec_send_frame.argtypes = [C.c_void_p, C.c_void_p]
    av.avcodec_receive_packet.argtypes = [C.c_void_p, C.c_void_p]
    av.av_new_packet.argtypes = [C.c_void_p, C.c_int]
    av.av_packet_unref.argtypes = [C.c_void_p]
    av.avcodec_free_context.argtypes = [C.c_void_p]
    u.av_frame_alloc.restype = C.c_void_p
    u.av_frame_unref.argtypes = [C.c_void_p]
    u.av_frame_get_buffer.argtypes = [C.c_void_p, C.c_int]
    u.av_opt_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    u.av_opt_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int]
    u.av_opt_get_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    u.av_channel_layout_default.argtypes = [C.c_void_p, C.c_int]
    u.av_log_set_level.argtypes = [C.c_int]
    u.av_strerror.argtypes = [C.c_int, C.c_char_p, C.c_size_t]
    return _libs


def _err(code):
    b = C.create_string_buffer(128)
    _load()["libavutil"].av_strerror(code, b, 128)
    return "%d %s" % (code, b.value.decode())


# AVPacket prefix (stable since FFmpeg 4): buf(8) pts(8) dts(8) data(8) size(4)
PKT_DATA, PKT_SIZE = 24, 32
# AVFrame prefix: data[8] @0, linesize[8] @64, extended_data @96, width @104, height @108, nb_samples @112, format @116
FR_DATA, FR_EXT, FR_NB, FR_FMT = 0, 96, 112, 116

SAMPLE_FMTS = {0: "u8", 1: "s16", 2: "s32", 3: "flt", 4: "dbl", 5: "u8p", 6: "s16p", 7: "s32p", 8: "fltp", 9: "dblp"}


def split_blocks(data):
    """Yield (offset, size, header fields) per 'wvpk' block by hopping ckSize+8."""
    off = 0
    while off + 32 <= len(data):
        if data[off:off + 4] != b"wvpk":
            raise ValueError("lost sync at %d" % off)
        cks, ver = struct.unpack_from("<IH", data, off + 4)
        total, index, nsamp, flags, crc = struct.unpack_from("<IIIII", data, off + 12)
        yield off, cks + 8, dict(version=ver, total=total, index=index, samples=nsamp, flags=flags, crc=crc)
        off += cks + 8


def packets(data):
    """Group blocks into decoder packets: INITIAL..FINAL blocks of one segment together."""
    cur = None
    for off, size, h in split_blocks(data):
        if cur is None:
            cur = [off, 0]
        cur[1] += size
        if h["flags"] & 0x1000 or h["samples"] == 0:
            yield data[cur[0]:cur[0] + cur[1]], h
            cur = None
    if cur is not None:
        yield data[cur[0]:cur[0] + cur[1]], None


def ff_decode(data, nch):
    """Decode a .wv byte string with FFmpeg's native decoder.
    Returns (interleaved int32/float32 numpy array shaped (n, nch), sample_fmt name)."""
    L = _load()
    av, u = L["libavcodec"], L["libavutil"]
    u.av_log_set_level(16)
    codec = av.avcodec_find_decoder_by_name(b"wavpack")
    ctx = av.avcodec_alloc_context3(codec)
    rc = av.avcodec_open2(ctx, codec, None)
    if rc < 0:
        raise RuntimeError("open2 " + _err(rc))
    pkt = av.av_packet_alloc()
    frame = u.av_frame_alloc()
    outs = []
    fmtname = None
    for chunk, h in packets(data):
        if h is not None and h["samples"] == 0:
            continue
        rc = av.av_new_packet(pkt, len(chunk))
        assert rc == 0
        dptr = C.c_void_p.from_address(pkt + PKT_DATA).value
        C.memmove(dptr, chunk, len(chunk))
        rc = av.avcodec_send_packet(ctx, pkt)
        av.av_packet_unref(pkt)
        if rc < 0:
            raise RuntimeError("send_packet " + _err(rc))
        while True:
            rc = av.avcodec_receive_frame(ctx, frame)
            if rc < 0:
                break
            nb = C.c_int.from_address(frame + FR_NB).value
            fmt = C.c_int.from_address(frame + FR_FMT).value
            fmtname = SAMPLE_FMTS.get(fmt, str(fmt))
            ext = C.c_void_p.from_address(frame + FR_EXT).value
            planes = []
            bps = {"u8p": 1, "s16p": 2, "s32p": 4, "fltp": 4}[fmtname]
            dt = {"u8p": np.uint8, "s16p": np.int16, "s32p": np.int32, "fltp": np.float32}[fmtname]
            for c in range(nch):
                p = C.c_void_p.from_address(ext + 8 * c).value
                planes.append(np.frombuffer(C.string_at(p, nb * bps), dtype=dt).copy())
            outs.append(np.stack(planes, axis=1))
            u.av_frame_unref(frame)
    av.avcodec_free_context(C.byref(C.c_void_p(ctx)))
    return (np.concatenate(outs, axis=0) if outs else np.zeros((0, nch))), fmtname


if __name__ == "__main__":
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
    from _harness import make_file
    cfg, src, data = make_file(seconds=1.0)
    out, fmt = ff_decode(data, 2)
    print(fmt, out.shape, np.array_equal(out.reshape(-1).astype(np.int32), src))


# ----------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------
_frame_offsets = None


def _discover_frame_offsets():
    """Find AVFrame.sample_rate / AVFrame.ch_layout offsets by decoding a known stream
    (44100 Hz stereo) and scanning the frame for the known values."""
    global _frame_offsets
    if _frame_offsets:
        return _frame_offsets
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from _harness import make_file
    cfg, src, data = make_file(seconds=0.05, block_samples=1000)
    L = _load()
    av, u = L["libavcodec"], L["libavutil"]
    codec = av.avcodec_find_decoder_by_name(b"wavpack")
    ctx = av.avcodec_alloc_context3(codec)
    assert av.avcodec_open2(ctx, codec, None) == 0
    pkt = av.av_packet_alloc()
    frame = u.av_frame_alloc()
    chunk, _ = next(packets(data))
    av.av_new_packet(pkt, len(chunk))
    C.memmove(C.c_void_p.from_address(pkt + PKT_DATA).value, chunk, len(chunk))
    assert av.avcodec_send_packet(ctx, pkt) == 0
    assert av.avcodec_receive_frame(ctx, frame) == 0
    craw = C.string_at(ctx, 1200)
    fmt_off = None
    for off in range(8, 1190, 4):
        a, b = struct.unpack_from("<ii", craw, off)
        if a == 44100 and b == 6:  # sample_rate followed by sample_fmt (s16p)
            fmt_off = off + 4
            break
    assert fmt_off, "AVCodecContext.sample_fmt not found"
    raw = C.string_at(frame, 512)
    sr_off = None
    for off in range(120, 500, 4):
        if struct.unpack_from("<i", raw, off)[0] == 44100:
            sr_off = off
            break
    cl_off = None
    for off in range(120, 480, 8):
        order, nb, mask = struct.unpack_from("<iiQ", raw, off)
        if order == 1 and nb == 2 and mask == 3:
            cl_off = off
    assert sr_off and cl_off, (sr_off, cl_off)
    u.av_frame_unref(frame)
    av.avcodec_free_context(C.byref(C.c_void_p(ctx)))
    _frame_offsets = (sr_off, cl_off, fmt_off)
    return _frame_offsets


def ff_encode(pcm, nch, rate, fmt="s16p", compression_level=None, opts=None):
    """Encode interleaved samples (numpy, shape (n*nch,) or (n,nch)) with FFmpeg's native WavPack
    encoder.  fmt: u8p/s16p/s32p/fltp.  Returns the concatenated packet bytes (= a raw .wv stream)."""
    L = _load()
    av, u = L["libavcodec"], L["libavutil"]
    u.av_log_set_level(16)
    sr_off, cl_off, ctx_fmt_off = _discover_frame_offsets()
    pcm = np.asarray(pcm).reshape(-1, nch)
    codec = av.avcodec_find_encoder_by_name(b"wavpack")
    ctx = av.avcodec_alloc_context3(codec)
    S = 1  # AV_OPT_SEARCH_CHILDREN

    def opt(k, v):
        rc = u.av_opt_set(ctx, k.encode(), str(v).encode(), S)
        if rc < 0:
            raise RuntimeError("av_opt_set %s=%s: %s" % (k, v, _err(rc)))

    opt("ar", rate)
    fmt_id = {v: k for k, v in SAMPLE_FMTS.items()}[fmt]
    C.c_int.from_address(ctx + ctx_fmt_off).value = fmt_id
    opt("ch_layout", {1: "mono", 2: "stereo", 6: "5.1"}[nch])
    opt("time_base", "1/%d" % rate)
    if compression_level is not None:
        opt("compression_level", compression_level)
    for k, v in (opts or {}).items():
        opt(k, v)
    rc = av.avcodec_open2(ctx, codec, None)
    if rc < 0:
        raise RuntimeError("open2 " + _err(rc))
    fs = C.c_int64()
    assert u.av_opt_get_int(ctx, b"frame_size", S, C.byref(fs)) == 0
    frame_size = fs.value or 4096
    fmt_id = {v: k for k, v in SAMPLE_FMTS.items()}[fmt]
    dt = {"u8p": np.uint8, "s16p": np.int16, "s32p": np.int32, "fltp": np.float32}[fmt]
    pkt = av.av_packet_alloc()
    frame = u.av_frame_alloc()
    out = bytearray()

    def drain():
        while True:
            rc = av.avcodec_receive_packet(ctx, pkt)
            if rc < 0:
                return
            p = C.c_void_p.from_address(pkt + PKT_DATA).value
            n = C.c_int.from_address(pkt + PKT_SIZE).value
            out.extend(C.string_at(p, n))
            av.av_packet_unref(pkt)

    pos = 0
    n = pcm.shape[0]
    while pos < n:
        k = min(frame_size, n - pos)
        C.c_int.from_address(frame + FR_NB).value = k
        C.c_int.from_address(frame + FR_FMT).value = fmt_id
        C.c_int.from_address(frame + sr_off).value = rate
        u.av_channel_layout_default(frame + cl_off, nch)
        rc = u.av_frame_get_buffer(frame, 0)
        if rc < 0:
            raise RuntimeError("get_buffer " + _err(rc))
        ext = C.c_void_p.from_address(frame + FR_EXT).value
        for c in range(nch):
            plane = np.ascontiguousarray(pcm[pos:pos + k, c].astype(dt))
            C.memmove(C.c_void_p.from_address(ext + 8 * c).value, plane.ctypes.data, plane.nbytes)
        rc = av.avcodec_send_frame(ctx, frame)
        if rc < 0:
            raise RuntimeError("send_frame " + _err(rc))
        u.av_frame_unref(frame)
        drain()
        pos += k
    av.avcodec_send_frame(ctx, None)
    drain()
    av.avcodec_free_context(C.byref(C.c_void_p(ctx)))
    return bytes(out)
