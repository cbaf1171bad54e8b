"""Test/bench helpers: ctypes bindings for the oracle (oracle/librefdec.so) and the
synthetic corpus generator (corpus/libwvenc.so).  Test infrastructure only -- the
product package never imports this module."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(subdir, lib):
    path = os.path.join(ROOT, subdir, lib)
    srcs = [os.path.join(ROOT, subdir, f) for f in os.listdir(os.path.join(ROOT, subdir)) if f.endswith((".c", ".h"))]
    if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, subdir)])
    return path


# ----------------------------------------------------------------------------
# corpus generator
# ----------------------------------------------------------------------------
class WvencConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("sample_rate", C.c_int32), ("bits", C.c_int32), ("channels", C.c_int32),
        ("block_samples", C.c_int32), ("nterms", C.c_int32), ("terms", C.c_int8 * 16), ("deltas", C.c_int8 * 16),
        ("joint_stereo", C.c_int32), ("false_stereo", C.c_int32), ("shift", C.c_int32),
        ("int32_sent_bits", C.c_int32), ("int32_wvx", C.c_int32), ("int32_new_wvx", C.c_int32),
        ("int32_max_width", C.c_int32), ("int32_zeros", C.c_int32), ("int32_ones", C.c_int32), ("int32_dups", C.c_int32),
        ("hybrid_bitrate", C.c_int32), ("hybrid_balance", C.c_int32),
        ("float_flags", C.c_int32), ("float_shift", C.c_int32), ("float_max_exp", C.c_int32), ("float_norm_exp", C.c_int32),
        ("float_new_wvx", C.c_int32),
        ("dsd_mode", C.c_int32), ("dsd_rate_shift", C.c_int32), ("dsd_history_bits", C.c_int32),
        ("dsd_raw_probs", C.c_int32), ("dsd_rate_i", C.c_int32),
        ("extras", C.c_int32), ("version", C.c_int32), ("unknown_length", C.c_int32),
    ]


KIND_PCM, KIND_HYBRID, KIND_FLOAT, KIND_DSD = 0, 1, 2, 3
X_RIFF_HEADER, X_CONFIG, X_NEW_CONFIG, X_BLOCK_CHECKSUM, X_MD5_TRAILER, X_SAMPLE_RATE, X_DUMMY, X_ALL_HISTORY = (
    1, 2, 4, 8, 16, 32, 64, 128)

_wvenc = None


def wvenc():
    global _wvenc
    if _wvenc is None:
        lib = C.CDLL(_build("corpus", "libwvenc.so"))
        lib.wvenc_default_config.argtypes = [C.POINTER(WvencConfig)]
        lib.wvenc_synth.argtypes = [C.POINTER(WvencConfig), C.c_uint64, C.c_int64, C.c_void_p]
        lib.wvenc_encode.argtypes = [C.POINTER(WvencConfig), C.c_void_p, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.wvenc_encode.restype = C.c_size_t
        lib.wvenc_bound.argtypes = [C.POINTER(WvencConfig), C.c_int64]
        lib.wvenc_bound.restype = C.c_size_t
        lib.wvenc_build_corpus.argtypes = [C.POINTER(WvencConfig), C.c_int64, C.c_int64, C.c_uint64, C.c_int,
                                           C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.wvenc_build_corpus.restype = C.c_size_t
        lib.wvenc_dbg_table.argtypes = [C.c_int, C.c_int]
        _wvenc = lib
    return _wvenc


def make_config(**kw):
    cfg = WvencConfig()
    wvenc().wvenc_default_config(C.byref(cfg))
    terms = kw.pop("terms", None)
    deltas = kw.pop("deltas", None)
    if terms is not None:
        cfg.nterms = len(terms)
        for i, t in enumerate(terms):
            cfg.terms[i] = t
            cfg.deltas[i] = 2 if deltas is None else deltas[i]
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise KeyError(k)
        setattr(cfg, k, v)
    return cfg


def synth(cfg, seed, nsamples):
    out = np.zeros(nsamples * cfg.channels, dtype=np.int32)
    wvenc().wvenc_synth(C.byref(cfg), seed, nsamples, out.ctypes.data)
    return out


def encode(cfg, samples, want_recon=False):
    samples = np.ascontiguousarray(samples, dtype=np.int32)
    n = samples.size // cfg.channels
    cap = wvenc().wvenc_bound(C.byref(cfg), n)
    buf = np.zeros(cap, dtype=np.uint8)
    recon = np.zeros(samples.size, dtype=np.int32) if want_recon else None
    ln = wvenc().wvenc_encode(C.byref(cfg), samples.ctypes.data, n, buf.ctypes.data, cap,
                              recon.ctypes.data if want_recon else None)
    if ln == 0:
        raise RuntimeError("wvenc_encode failed")
    data = buf[:ln].tobytes()
    return (data, recon) if want_recon else data


def make_file(seed=0x5EED0000, seconds=1.0, want_recon=False, nsamples=None, **kw):
    cfg = make_config(**kw)
    n = int(cfg.sample_rate * seconds) if nsamples is None else nsamples
    if cfg.kind == KIND_DSD and nsamples is None:
        n = int(cfg.sample_rate * 8 * seconds)  # DSD64: 352800 byte-times/s at 44.1k base
    src = synth(cfg, seed, n)
    out = encode(cfg, src, want_recon)
    if want_recon:
        return cfg, src, out[0], out[1]
    return cfg, src, out


# ----------------------------------------------------------------------------
# oracle
# ----------------------------------------------------------------------------
_ref = None


def refdec():
    global _ref
    if _ref is None:
        lib = C.CDLL(_build("oracle", "librefdec.so"))
        lib.rd_open.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
        lib.rd_open.restype = C.c_void_p
        lib.rd_close.argtypes = [C.c_void_p]
        lib.rd_unpack_samples.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long]
        lib.rd_unpack_samples.restype = C.c_long
        lib.rd_format_samples.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_int, C.c_int]
        for name, res in [("rd_get_num_samples", C.c_long), ("rd_get_sample_index", C.c_long), ("rd_get_num_errors", C.c_long),
                          ("rd_lossy", C.c_int), ("rd_get_sample_rate", C.c_long), ("rd_get_num_channels", C.c_int),
                          ("rd_get_bits_per_sample", C.c_int), ("rd_get_bytes_per_sample", C.c_int),
                          ("rd_get_reduced_channels", C.c_int), ("rd_get_file_format", C.c_int),
                          ("rd_get_file_extension", C.c_char_p), ("rd_get_error_message", C.c_char_p),
                          ("rd_get_is_five", C.c_int), ("rd_get_version", C.c_int), ("rd_get_is_float", C.c_int),
                          ("rd_get_mode", C.c_int), ("rd_dbg_block_crc", C.c_int32), ("rd_dbg_mute_error", C.c_int),
                          ("rd_dbg_block_flags", C.c_uint32), ("rd_dbg_check_crc_error", C.c_int)]:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = [C.c_void_p] if name != "rd_get_num_samples" else [C.c_void_p, C.c_int]
        lib.rd_get_header.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        lib.rd_get_header.restype = C.c_void_p
        lib.rd_get_trailer.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        lib.rd_get_trailer.restype = C.c_void_p
        lib.rd_get_compression_level.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        lib.rd_dbg_unpack_current_block.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_long]
        lib.rd_dbg_unpack_current_block.restype = C.c_long
        lib.rd_decode_file_pcm.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_long, C.c_void_p, C.c_size_t,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_long)]
        lib.rd_decode_file_pcm.restype = C.c_long
        lib.rd_wvdemo.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        lib.rd_wvdemo.restype = C.c_long
        lib.rd_dbg_table.argtypes = [C.c_int, C.c_int]
        _ref = lib
    return _ref


class OracleFile:
    """WavpackContext of the oracle over an in-memory file."""

    def __init__(self, data, flags=0):
        self.lib = refdec()
        self._buf = np.frombuffer(data, dtype=np.uint8).copy()
        self.ctx = self.lib.rd_open(self._buf.ctypes.data, self._buf.size, flags)

    def close(self):
        if self.ctx:
            self.lib.rd_close(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    @property
    def error(self):
        m = self.lib.rd_get_error_message(self.ctx)
        return m.decode() if m else None

    def info(self):
        L = self.lib
        lvl = C.create_string_buffer(64)
        L.rd_get_compression_level(self.ctx, lvl, 64)
        hl, tl = C.c_long(), C.c_long()
        hp = L.rd_get_header(self.ctx, C.byref(hl))
        tp = L.rd_get_trailer(self.ctx, C.byref(tl))
        return dict(
            num_samples=L.rd_get_num_samples(self.ctx, 0), num_samples_native=L.rd_get_num_samples(self.ctx, 1),
            sample_rate=L.rd_get_sample_rate(self.ctx), num_channels=L.rd_get_num_channels(self.ctx),
            reduced_channels=L.rd_get_reduced_channels(self.ctx), bits_per_sample=L.rd_get_bits_per_sample(self.ctx),
            bytes_per_sample=L.rd_get_bytes_per_sample(self.ctx), lossy=bool(L.rd_lossy(self.ctx)),
            file_format=L.rd_get_file_format(self.ctx), file_extension=L.rd_get_file_extension(self.ctx).decode("utf-8", "replace"),
            is_five=bool(L.rd_get_is_five(self.ctx)), version=L.rd_get_version(self.ctx),
            is_float=bool(L.rd_get_is_float(self.ctx)), mode=L.rd_get_mode(self.ctx),
            compression_level=lvl.value.decode() or None,
            header=C.string_at(hp, hl.value) if hp else None, trailer=C.string_at(tp, tl.value) if tp else None,
        )

    def unpack(self, samples, nch=None):
        nch = nch or self.lib.rd_get_reduced_channels(self.ctx)
        buf = np.zeros(samples * nch, dtype=np.int32)
        n = self.lib.rd_unpack_samples(self.ctx, buf.ctypes.data, buf.size, samples)
        return n, buf

    def decode_all(self, chunk=4096):
        """WvDemo-style loop; returns (int32 interleaved array, crc_errors, status)."""
        nch = self.lib.rd_get_reduced_channels(self.ctx)
        out = []
        status = 0
        while True:
            n, buf = self.unpack(chunk, nch)
            if n < 0:
                status = n
                break
            if n == 0:
                break
            out.append(buf[: n * nch].copy())
        data = np.concatenate(out) if out else np.zeros(0, dtype=np.int32)
        return data, self.lib.rd_get_num_errors(self.ctx), status


def oracle_decode(data, flags=0, chunk=4096):
    f = OracleFile(data, flags)
    if f.error:
        err = f.error
        f.close()
        raise RuntimeError(err)
    out, errs, status = f.decode_all(chunk)
    info = f.info()
    info["lossy"] = bool(f.lib.rd_lossy(f.ctx))  # lossy_blocks is updated while decoding
    info["mode"] = f.lib.rd_get_mode(f.ctx)
    f.close()
    return out, errs, status, info


def format_samples(src, bps, dsd=False):
    """WavpackFormatSamples through the oracle."""
    src = np.ascontiguousarray(src, dtype=np.int32)
    pcm = np.zeros(src.size * bps, dtype=np.uint8)
    ok = refdec().rd_format_samples(src.ctypes.data, src.size, bps, pcm.ctypes.data, pcm.size, 0, int(dsd))
    assert ok
    return pcm


def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()


# ----------------------------------------------------------------------------
# device-code emulation (tests/emul): the CUDA per-thread decode function compiled for the host.
# Development/regression aid for boxes without a GPU; never used by the product.
# ----------------------------------------------------------------------------
_emul = None


def emul():
    global _emul
    if _emul is None:
        from wavpackdecoder_b200 import _native as N
        src = [os.path.join(ROOT, "tests", "emul", "emul.cpp"), os.path.join(ROOT, "wavpackdecoder_b200", "csrc", "wvb_index.cpp")]
        deps = src + [os.path.join(ROOT, "wavpackdecoder_b200", "csrc", f) for f in ("wvb_pcm.cuh", "wvb_dsd.cuh", "wvb_plan.h", "wv_tables.h")]
        deps = [d for d in deps if os.path.exists(d)] + [os.path.join(ROOT, "include", "wvb.h")]
        path = os.path.join(ROOT, "tests", "emul", "libwvb_emul.so")
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-fwrapv", "-Wno-unknown-pragmas", "-shared", "-pthread",
                                   "-o", path] + src)
        lib = C.CDLL(path)
        N.declare_index_api(lib)
        lib.emul_decode.argtypes = [C.c_void_p, C.POINTER(N.BlockDesc), C.c_size_t, C.c_void_p, C.c_int, C.POINTER(N.BlockResult)]
        _emul = lib
    return _emul


def index_file(lib, data, open_flags=0, chunk=4096):
    """wvb_index over one in-memory file -> (FileInfo, ctypes array of BlockDesc)."""
    from wavpackdecoder_b200 import _native as N
    buf = np.frombuffer(data, dtype=np.uint8)
    info = N.FileInfo()
    n = C.c_size_t()
    rc = lib.wvb_index(buf.ctypes.data, buf.size, open_flags, chunk, C.byref(info), None, 0, C.byref(n))
    assert rc == 0, rc
    descs = (N.BlockDesc * max(n.value, 1))()
    info = N.FileInfo()
    rc = lib.wvb_index(buf.ctypes.data, buf.size, open_flags, chunk, C.byref(info), descs, n.value, C.byref(n))
    assert rc == 0, rc
    return info, descs, n.value


def emul_decode_file(data, open_flags=0, chunk=4096, out_format=0):
    """Index + decode one file through the host-compiled device code.
    Returns (numpy output [int32 or uint8], info, results list)."""
    from wavpackdecoder_b200 import _native as N
    lib = emul()
    info, descs, n = index_file(lib, data, open_flags, chunk)
    if info.status != 0:
        raise RuntimeError(info.error_message.decode())
    nch = info.num_channels if (open_flags & N.OPEN_ALL_CHANNELS) else (info.reduced_channels or info.num_channels)
    unit = 4 if out_format == 0 else info.bytes_per_sample
    lib.wvb_rebase(descs, n, 0, 0, out_format, 0)
    out = np.zeros(info.indexed_samples * nch * unit + 64, dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8)
    padded = np.concatenate([buf, np.zeros(64, dtype=np.uint8)])
    res = (N.BlockResult * max(n, 1))()
    lib.emul_decode(padded.ctypes.data, descs, n, out.ctypes.data, out_format, res)
    out = out[: info.indexed_samples * nch * unit]
    if out_format == 0:
        out = out.view(np.int32)
    return out, info, [res[i] for i in range(n)], [descs[i] for i in range(n)]


def oracle_wvdemo(data):
    """WvDemo.Main restated by the oracle: (output file bytes, exit code)."""
    lib = refdec()
    buf = np.frombuffer(data, dtype=np.uint8)
    cap = 64 + buf.size * 40 + (1 << 16)
    while True:
        out = np.zeros(cap, dtype=np.uint8)
        code = C.c_int(-1)
        n = lib.rd_wvdemo(buf.ctypes.data, buf.size, out.ctypes.data, cap, C.byref(code))
        if n >= 0:
            return out[:n].tobytes(), code.value
        cap *= 4


class EmulDecoder:
    """Stand-in for wavpackdecoder_b200.batch.BatchDecoder backed by the host-compiled device code (tests/emul).
    TEST SEAM ONLY: CPU tests of the API mirror's window / seek logic install it as `wpc._dec`; the product never does."""

    def __init__(self):
        self.lib = emul()

    def decode(self, in_ptr, in_bytes, descs, nblocks, out_ptr, out_bytes, out_format, mem_flags=0, results=None):
        assert mem_flags == 0
        self.lib.emul_decode(in_ptr, descs, nblocks, out_ptr, out_format, results)

    def close(self):
        pass
