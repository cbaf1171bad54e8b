"""ctypes binding of the C ABI in include/wvb.h (libwvb.so).

This module only declares structures and prototypes; it adds no behaviour.  Loading fails
loudly if the library has not been built (``python -c "import __graft_entry__ as g; g.build()"``)
-- there is no Python or CPU fallback for the decode path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WVB_LIB") or os.path.join(HERE, "libwvb.so")  # WVB_LIB: alternative build for tuning experiments

WVB_SUB_COUNT = 8
SUB_TERMS, SUB_WEIGHTS, SUB_SAMPLES, SUB_ENTROPY, SUB_HYBRID, SUB_WV, SUB_WVX, SUB_DSD = range(8)

OK, E_ARG, E_NO_DEVICE, E_CUDA, E_CAPACITY, E_FORMAT = 0, -1, -2, -3, -4, -5
OPEN_2CH_MAX = 0x8
OPEN_ALL_CHANNELS = 0x10000
OUT_INT32, OUT_PCM, OUT_DSD_RAW = 0, 1, 2
IN_DEVICE, OUT_DEVICE, RESULTS_DEVICE, NO_SYNC = 1, 2, 4, 8
RF_CRC_ERROR, RF_MUTED, RF_CRCX_ERROR, RF_INEXACT, RF_BAD_BLOCK, RF_BLOCK_CHECKSUM = 1, 2, 4, 8, 16, 32
BF_WVX_NEW, BF_HAS_INT32_INFO, BF_HAS_FLOAT_INFO, BF_WVX_PRESENT, BF_MUTE_ALL, BF_STALE_STATE, BF_DSD_PADDED, BF_BLOCK_CHECKSUM = 1, 2, 4, 8, 16, 32, 64, 128


class BlockDesc(C.Structure):
    _fields_ = [
        ("in_offset", C.c_uint64), ("out_offset", C.c_uint64), ("in_bytes", C.c_uint32), ("block_samples", C.c_uint32),
        ("flags", C.c_uint32), ("crc", C.c_int32), ("block_index", C.c_int64),
        ("sub_off", C.c_uint32 * WVB_SUB_COUNT), ("sub_len", C.c_uint32 * WVB_SUB_COUNT),
        ("int32_info", C.c_uint8 * 4), ("float_info", C.c_uint8 * 4), ("bflags", C.c_uint32), ("version", C.c_uint16),
        ("out_channels", C.c_uint8), ("out_stride", C.c_uint8), ("out_ch_offset", C.c_uint8), ("out_bps", C.c_uint8),
        ("smem_words", C.c_uint16), ("chunk_first", C.c_uint32), ("chunk_samples", C.c_uint32), ("file_id", C.c_uint32),
        ("gap_before", C.c_uint32), ("terms_sig", C.c_uint32), ("skip_samples", C.c_uint32), ("skip_chunk", C.c_uint32),
        ("avg_block_size", C.c_uint32), ("checksum_off", C.c_uint32),
    ]


class SeekState(C.Structure):
    _fields_ = [("hdr_pos", C.c_int64), ("block_index", C.c_int64), ("avg_block_size", C.c_int64), ("file_pos", C.c_int64),
                ("block_samples", C.c_uint32), ("ck_size", C.c_uint32)]


class BlockResult(C.Structure):
    _fields_ = [("crc", C.c_int32), ("rflags", C.c_uint32), ("mute_from", C.c_uint32), ("crc_x", C.c_int32)]


class FileInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("error_message", C.c_char * 64), ("total_samples", C.c_int64), ("sample_rate", C.c_int64),
        ("config_flags", C.c_int64), ("channel_mask", C.c_int64), ("num_channels", C.c_int32), ("reduced_channels", C.c_int32),
        ("bits_per_sample", C.c_int32), ("bytes_per_sample", C.c_int32), ("float_norm_exp", C.c_int32), ("xmode", C.c_int32),
        ("version", C.c_int32), ("five", C.c_int32), ("file_format", C.c_int32), ("lossy_blocks", C.c_int32),
        ("dsd_multiplier", C.c_uint32), ("first_flags", C.c_uint32), ("header_off", C.c_int64), ("header_len", C.c_int64),
        ("trailer_off", C.c_int64), ("trailer_len", C.c_int64), ("file_extension", C.c_char * 16), ("num_blocks", C.c_int64),
        ("indexed_samples", C.c_int64), ("stopped_early", C.c_int32), ("reserved", C.c_int32),
    ]


assert C.sizeof(BlockDesc) == 160, C.sizeof(BlockDesc)


def desc_table(descs, n):
    """numpy structured view (no copy) of a ctypes BlockDesc array: host-side bookkeeping over 10^5 descriptors
    (shard costs, per-file error counts) without a Python loop."""
    import numpy as np
    dt = np.dtype([("in_offset", "<u8"), ("out_offset", "<u8"), ("in_bytes", "<u4"), ("block_samples", "<u4"), ("flags", "<u4"), ("crc", "<i4"),
                   ("block_index", "<i8"), ("sub_off", "<u4", (8,)), ("sub_len", "<u4", (8,)), ("int32_info", "u1", (4,)), ("float_info", "u1", (4,)),
                   ("bflags", "<u4"), ("version", "<u2"), ("out_channels", "u1"), ("out_stride", "u1"), ("out_ch_offset", "u1"), ("out_bps", "u1"),
                   ("smem_words", "<u2"), ("chunk_first", "<u4"), ("chunk_samples", "<u4"), ("file_id", "<u4"), ("gap_before", "<u4"),
                   ("terms_sig", "<u4"), ("skip_samples", "<u4"), ("skip_chunk", "<u4"), ("avg_block_size", "<u4"), ("checksum_off", "<u4")])
    assert dt.itemsize == C.sizeof(BlockDesc)
    return np.frombuffer(descs, dtype=dt, count=n)


def result_table(results, n):
    import numpy as np
    dt = np.dtype([("crc", "<i4"), ("rflags", "<u4"), ("mute_from", "<u4"), ("crc_x", "<i4")])
    return np.frombuffer(results, dtype=dt, count=n)
assert C.sizeof(BlockResult) == 16

# every symbol include/wvb.h declares; tests/test_abi.py checks the built library exports all of them
EXPORTS = [
    "wvb_abi_version", "wvb_abi_layout", "wvb_last_error", "wvb_device_count", "wvb_index", "wvb_index_seek", "wvb_index_many", "wvb_rebase", "wvb_frame_bytes",
    "wvb_batch_create", "wvb_batch_destroy", "wvb_batch_prepare", "wvb_batch_decode", "wvb_batch_wait", "wvb_batch_timing", "wvb_batch_stream",
    "wvb_host_alloc", "wvb_host_free", "wvb_batch_md5", "wvb_stored_md5", "wvb_block_checksum_ok", "wvb_batch_decode_files", "wvb_batch_dsd_to_dsf",
]


def declare_index_api(lib):
    """Prototypes of the host-only part of the ABI (also exported by the test emulation library)."""
    lib.wvb_index.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.POINTER(FileInfo), C.POINTER(BlockDesc),
                              C.c_size_t, C.POINTER(C.c_size_t)]
    lib.wvb_index.restype = C.c_int
    lib.wvb_index_seek.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(SeekState), C.c_int64, C.c_uint32, C.c_uint32, C.c_size_t,
                                   C.POINTER(FileInfo), C.POINTER(BlockDesc), C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64)]
    lib.wvb_index_seek.restype = C.c_int
    lib.wvb_index_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                   C.POINTER(FileInfo), C.POINTER(BlockDesc), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
    lib.wvb_index_many.restype = C.c_int
    lib.wvb_rebase.argtypes = [C.POINTER(BlockDesc), C.c_size_t, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32]
    lib.wvb_rebase.restype = None
    lib.wvb_frame_bytes.argtypes = [C.POINTER(BlockDesc), C.c_int]
    lib.wvb_frame_bytes.restype = C.c_uint32
    lib.wvb_stored_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    lib.wvb_stored_md5.restype = C.c_int
    lib.wvb_abi_layout.restype = C.c_char_p
    lib.wvb_block_checksum_ok.argtypes = [C.c_void_p, C.c_size_t]
    lib.wvb_block_checksum_ok.restype = C.c_int
    return lib


def check_layout(lib):
    """The ctypes mirrors above against the layouts the library was compiled with (wvb_abi_layout)."""
    mirrors = {"wvb_block_desc": BlockDesc, "wvb_block_result": BlockResult, "wvb_file_info": FileInfo, "wvb_seek_state": SeekState}
    for part in lib.wvb_abi_layout().decode().split("|"):
        items = [x for x in part.split(";") if x]
        name, size = items[0].split(":")
        cls = mirrors[name]
        if C.sizeof(cls) != int(size):
            raise RuntimeError("%s: ctypes mirror is %d bytes, library says %s" % (name, C.sizeof(cls), size))
        for it in items[1:]:
            f, off, sz = it.split(":")
            d = getattr(cls, f)
            if d.offset != int(off) or d.size != int(sz):
                raise RuntimeError("%s.%s: ctypes mirror at %d (%d bytes), library says %s (%s bytes)" % (name, f, d.offset, d.size, off, sz))


_lib = None


def load():
    """Load libwvb.so.  Raises if it is missing: the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "wavpackdecoder_b200: %s not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc). There is no CPU fallback for the decode path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    declare_index_api(lib)
    lib.wvb_abi_version.restype = C.c_int
    lib.wvb_last_error.restype = C.c_char_p
    lib.wvb_device_count.restype = C.c_int
    lib.wvb_batch_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.wvb_batch_create.restype = C.c_int
    lib.wvb_batch_destroy.argtypes = [C.c_void_p]
    lib.wvb_batch_destroy.restype = None
    lib.wvb_batch_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(BlockDesc), C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.c_int, C.c_uint32, C.c_void_p]
    lib.wvb_batch_decode.restype = C.c_int
    lib.wvb_batch_decode_files.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                           C.c_int, C.POINTER(FileInfo), C.POINTER(BlockDesc), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
    lib.wvb_batch_decode_files.restype = C.c_int
    lib.wvb_batch_dsd_to_dsf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.wvb_batch_dsd_to_dsf.restype = C.c_int
    lib.wvb_batch_prepare.argtypes = [C.c_void_p, C.POINTER(BlockDesc), C.c_size_t, C.c_int]
    lib.wvb_batch_prepare.restype = C.c_int
    lib.wvb_batch_wait.argtypes = [C.c_void_p]
    lib.wvb_batch_wait.restype = C.c_int
    lib.wvb_batch_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.wvb_batch_timing.restype = C.c_int
    lib.wvb_batch_stream.argtypes = [C.c_void_p]
    lib.wvb_batch_stream.restype = C.c_void_p
    lib.wvb_batch_md5.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.wvb_batch_md5.restype = C.c_int
    lib.wvb_host_alloc.argtypes = [C.c_size_t]
    lib.wvb_host_alloc.restype = C.c_void_p
    lib.wvb_host_free.argtypes = [C.c_void_p]
    lib.wvb_host_free.restype = None
    if lib.wvb_abi_version() != 3:
        raise RuntimeError("libwvb.so ABI version mismatch")
    check_layout(lib)
    _lib = lib
    return lib
