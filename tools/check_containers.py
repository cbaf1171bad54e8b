#!/usr/bin/env python3
"""Independent check of the synthesised container headers (wavpackdecoder_b200/containers.py): open them with the
demuxers of the FFmpeg libavformat that ships inside this container's opencv wheel (w64, caf, iff/DSDIFF, wav) and compare
the packet payload with the audio bytes that were put in, plus the sample rate / channel count the demuxer reports.
Offline tool (like tools/ffwv.py): nothing in tests/ or the product imports it.  python tools/check_containers.py"""
import ctypes as C
import glob
import os
import struct
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ffwv  # noqa: E402  (loads libavutil / libavcodec with RTLD_GLOBAL)
from wavpackdecoder_b200 import containers as K  # noqa: E402
from wavpackdecoder_b200.wvdemo import wave_header  # noqa: E402


def avformat():
    ffwv._load()
    hits = glob.glob(os.path.join(ffwv._LIBDIR, "libavformat-*"))
    lib = C.CDLL(hits[0], mode=C.RTLD_GLOBAL)
    lib.avformat_open_input.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_void_p, C.c_void_p]
    lib.avformat_find_stream_info.argtypes = [C.c_void_p, C.c_void_p]
    lib.av_read_frame.argtypes = [C.c_void_p, C.c_void_p]
    lib.avformat_close_input.argtypes = [C.POINTER(C.c_void_p)]
    lib.av_find_input_format.argtypes = [C.c_char_p]
    lib.av_find_input_format.restype = C.c_void_p
    return lib


def demux(lib, path, fmt_name):
    """(payload bytes, first 64 int32 of the stream's codec parameters)"""
    av = ffwv._libs["libavcodec"]
    ctx = C.c_void_p()
    fmt = lib.av_find_input_format(fmt_name.encode())
    assert fmt, "demuxer %s not in this libavformat" % fmt_name
    rc = lib.avformat_open_input(C.byref(ctx), path.encode(), fmt, None)
    assert rc == 0, "avformat_open_input(%s) = %d" % (fmt_name, rc)
    assert lib.avformat_find_stream_info(ctx, None) >= 0
    streams = C.cast(C.c_void_p.from_address(ctx.value + 48).value, C.POINTER(C.c_void_p))  # AVFormatContext.streams
    codecpar = C.c_void_p.from_address(streams[0] + 16).value                               # AVStream.codecpar
    par = (C.c_int32 * 64).from_address(codecpar)
    av.av_packet_alloc.restype = C.c_void_p
    pkt = av.av_packet_alloc()
    out = bytearray()
    while lib.av_read_frame(ctx, pkt) >= 0:
        data = C.c_void_p.from_address(pkt + 24).value
        size = C.c_int.from_address(pkt + 32).value
        out += C.string_at(data, size)
        av.av_packet_unref(pkt)
    pars = list(par)
    lib.avformat_close_input(C.byref(ctx))
    return bytes(out), pars


def main():
    lib = avformat()
    ok = True
    pcm = bytes((i * 37 + (i >> 8)) & 0xff for i in range(44100 * 2 * 2 // 10))  # 0.1 s of 16-bit stereo
    n = len(pcm) // 4
    dsd = bytes((i * 91) & 0xff for i in range(352800 * 2 // 100))            # 10 ms of DSD64 stereo
    nd = len(dsd) // 2
    cases = [
        ("wav", "wav", wave_header(n, 2, 44100, 16, 2) + pcm, pcm, 44100, 2),
        ("w64", "w64", K.w64_header(n, 2, 44100, 16, 2) + pcm + K.w64_trailer(n, 2, 2), pcm, 44100, 2),
        ("caf", "caf", K.caf_header(n, 2, 44100, 16, 2) + pcm, pcm, 44100, 2),
        ("w64 24-bit mono", "w64", K.w64_header(999, 1, 48000, 24, 3) + pcm[:2997] + K.w64_trailer(999, 1, 3), pcm[:2997], 48000, 1),
        ("dff", "iff", K.dff_header(nd, 2, 2822400) + dsd + K.dff_trailer(nd, 2), dsd, 2822400 // 8, 2),
    ]
    # DSF: the data chunk holds the re-laid-out bytes (wvb_batch_dsd_to_dsf on the device; here tests/test_containers.py's numpy
    # restatement of the layout), which FFmpeg's dsf demuxer hands out block by block as they are stored
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_containers import dsf_layout
    dsf_data = dsf_layout(dsd, nd, 2)
    # (FFmpeg's demuxer drops the zero padding of the last block group: each channel's valid bytes, channel after channel)
    groups, valid = (nd + 4095) // 4096, nd - (nd + 4095) // 4096 * 4096 + 4096
    expect = dsf_data[:(groups - 1) * 8192] + b"".join(dsf_data[(groups - 1) * 8192 + c * 4096:(groups - 1) * 8192 + c * 4096 + valid] for c in range(2))
    cases.append(("dsf", "dsf", K.dsf_header(nd, 2, 2822400) + dsf_data, expect, 2822400 // 8, 2))
    for name, demuxer, blob, audio, rate, ch in cases:
        with tempfile.NamedTemporaryFile(suffix="." + name.split()[0], delete=False) as f:
            f.write(blob)
            path = f.name
        try:
            payload, pars = demux(lib, path, demuxer)
        finally:
            os.unlink(path)
        same = payload == audio
        # (AVCodecParameters' layout differs between FFmpeg versions: look for the values rather than fixed offsets;
        #  FFmpeg reports DSD streams at the byte rate, 1/8 of the one-bit rate in the FS chunk)
        has_rate, has_ch = rate in pars, ch in pars
        print("%-16s demuxer %-4s payload %s (%d bytes)  sample rate %s  channels %s" % (
            name, demuxer, "identical" if same else "DIFFERS", len(payload), "found" if has_rate else "NOT found", "found" if has_ch else "NOT found"))
        ok = ok and same and has_rate and has_ch
    print("all containers accepted by FFmpeg's demuxers" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
