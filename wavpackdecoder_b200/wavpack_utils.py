"""Host-side mirror of the reference's public API (class WavPack.WavPackUtils, WavPackUtils.cs:14-672) over libwvb.

Same names, argument meaning and error behaviour as the C# original, so existing callers (WvDemo-style loops) port
one to one:

    wpc = WavpackOpenFileInput(data)                       # WavPackUtils.cs:36
    n = WavpackUnpackSamples(wpc, buffer, samples)         # WavPackUtils.cs:200  (buffer: numpy int32 array)
    WavpackFormatSamples(buffer, n * ch, bps, pcm)         # WavPackUtils.cs:288
    WavpackGetNumSamples / GetSampleRate / ... getters     # WavPackUtils.cs:133-499

Differences that are deliberate:
  * the stream is an in-memory bytes-like object (the reference takes a BinaryReader): the batch decoder needs the
    whole compressed file to ship it to the GPU;
  * the first WavpackUnpackSamples call decodes the WHOLE file on the device (every block in parallel) and later calls
    copy out of that result; the chunk size of the first call fixes the reference's chunk-dependent behaviour on
    corrupt streams (muting granularity), exactly as if every call used that size;
  * there is no CPU fallback: without libwvb.so / a CUDA device the unpack call raises.
No decode arithmetic lives here: Python only slices buffers and forwards getters.
"""
import ctypes as C

import numpy as np

from . import _native as N

SAMPLE_BUFFER_SIZE = 4096  # Defines.cs:18
OPEN_2CH_MAX = 0x8  # Defines.cs:26

# Defines.cs:137-147
MODE_WVC, MODE_LOSSLESS, MODE_HYBRID, MODE_FLOAT, MODE_VALID_TAG, MODE_HIGH, MODE_FAST, MODE_EXTRA = 1, 2, 4, 8, 0x10, 0x20, 0x40, 0x80
MODE_VERY_HIGH, MODE_XMODE, MODE_DSD = 0x400, 0x7000, 0x10000
CONFIG_HYBRID_FLAG, CONFIG_FLOAT_DATA, CONFIG_FAST_FLAG, CONFIG_HIGH_FLAG = 8, 0x80, 0x200, 0x800
CONFIG_VERY_HIGH_FLAG, CONFIG_LOSSY_MODE, CONFIG_EXTRA_MODE = 0x1000, 0x1000000, 0x2000000

FILE_FORMATS = ["WAV", "W64", "CAF", "DFF", "DSF", "AIF"]  # eFileFormat, Defines.cs:148-156


class WavpackContext:
    """Opaque context (WavpackContext.cs:13-36)."""

    def __init__(self):
        self.data = None
        self.info = None
        self.descs = None
        self.nblocks = 0
        self.error_message = None
        self.crc_errors = 0
        self.sample_index = 0
        self.open_flags = 0
        self.device = 0
        self._decoded = None      # int32 samples of the whole file, interleaved
        self._results = None
        self._block_ends = None   # cumulative sample positions at which each block's CRC verdict becomes visible
        self._chunk = None
        self._out_channels = 0


def WavpackOpenFileInput(infile, flags=0, device=0):
    """WavPackUtils.cs:36-120.  `infile`: bytes-like with the .wv stream.  Never raises for format errors: check
    WavpackGetErrorMessage(), like the reference."""
    lib = N.load()
    wpc = WavpackContext()
    wpc.data = np.frombuffer(infile, dtype=np.uint8)
    wpc.open_flags = flags
    wpc.device = device
    info = N.FileInfo()
    n = C.c_size_t()
    rc = lib.wvb_index(wpc.data.ctypes.data, wpc.data.size, flags, SAMPLE_BUFFER_SIZE, C.byref(info), None, 0, C.byref(n))
    if rc != N.OK:
        raise RuntimeError("wvb_index failed: %d" % rc)
    wpc.info = info
    if info.status != N.OK:
        wpc.error_message = info.error_message.decode() or "not compatible with this version of WavPack file!"
    wpc.sample_index = 0
    wpc._out_channels = info.reduced_channels or info.num_channels
    return wpc


def _decode_all(wpc, chunk):
    """Index with the caller's chunk size and decode every block on the device (int32 output)."""
    lib = N.load()
    from .batch import BatchDecoder
    info = N.FileInfo()
    n = C.c_size_t()
    lib.wvb_index(wpc.data.ctypes.data, wpc.data.size, wpc.open_flags, chunk, C.byref(info), None, 0, C.byref(n))
    descs = (N.BlockDesc * max(n.value, 1))()
    lib.wvb_index(wpc.data.ctypes.data, wpc.data.size, wpc.open_flags, chunk, C.byref(info), descs, n.value, C.byref(n))
    nblocks = n.value
    lib.wvb_rebase(descs, nblocks, 0, 0, N.OUT_INT32, 0)
    nch = wpc._out_channels
    out = np.zeros(int(info.indexed_samples) * nch + 16, dtype=np.int32)
    results = (N.BlockResult * max(nblocks, 1))()
    slab = np.concatenate([wpc.data, np.zeros(64, dtype=np.uint8)])
    dec = BatchDecoder(wpc.device)  # raises without a CUDA device: no fallback
    try:
        dec.decode(slab.ctypes.data, slab.size, descs, nblocks, out.ctypes.data, int(info.indexed_samples) * nch * 4, N.OUT_INT32, 0, results)
    finally:
        dec.close()
    wpc._decoded = out[: int(info.indexed_samples) * nch]
    wpc._results = [results[i] for i in range(nblocks)]
    wpc._block_ends = [int(descs[i].out_offset // (4 * nch)) + int(descs[i].block_samples) for i in range(nblocks)]
    wpc.descs, wpc.nblocks, wpc._chunk = descs, nblocks, chunk
    wpc.info.lossy_blocks = info.lossy_blocks


def WavpackUnpackSamples(wpc, buffer, samples):
    """WavPackUtils.cs:200-282: fills `buffer` (numpy int32, >= samples * channels) with right-justified samples and
    returns the number of complete samples unpacked (short at end of stream)."""
    if wpc.error_message:
        return 0
    if wpc._decoded is None:
        _decode_all(wpc, int(samples))
    nch = wpc._out_channels
    total = wpc._decoded.size // nch
    n = int(min(samples, total - wpc.sample_index))
    if wpc.info.total_samples >= 0 and wpc.sample_index < wpc.info.total_samples:
        n = int(min(n, wpc.info.total_samples - wpc.sample_index))  # the call returns at total_samples (WavPackUtils.cs:277)
    if n <= 0:
        return 0
    a = wpc.sample_index * nch
    buffer[: n * nch] = wpc._decoded[a: a + n * nch]
    new_index = wpc.sample_index + n
    # crc_errors becomes visible when the block's last sample has been handed out (WavPackUtils.cs:273-275)
    for end, r in zip(wpc._block_ends, wpc._results):
        if wpc.sample_index < end <= new_index and (r.rflags & N.RF_CRC_ERROR):
            wpc.crc_errors += 1
    wpc.sample_index = new_index
    return n


def WavpackFormatSamples(src, samcnt, bps, pcm_buffer, offset=0, dsd=False):
    """WavPackUtils.cs:288-341: int32 -> little-endian PCM of `bps` bytes; 8-bit gets +128 unless dsd.  Returns False when
    the destination is too small.  (The batch path produces packed PCM on the device with WVB_OUT_PCM; this helper exists
    for API parity with callers that format separately.)"""
    ln = int(samcnt) * bps
    if pcm_buffer is None or len(pcm_buffer) < ln + offset:
        return False
    v = np.asarray(src[:samcnt], dtype=np.int32)
    out = np.frombuffer(pcm_buffer, dtype=np.uint8) if not isinstance(pcm_buffer, np.ndarray) else pcm_buffer
    if bps == 1:
        out[offset:offset + ln] = (v if dsd else v + 128).astype(np.uint8)
    elif bps in (2, 3, 4):
        b = v.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :bps]
        out[offset:offset + ln] = b.reshape(-1)
    return True


def WavpackGetNumSamples(wpc, native=False):  # WavPackUtils.cs:346-350
    t = int(wpc.info.total_samples)
    return t * 8 if native and wpc.info.dsd_multiplier > 0 else t


def WavpackGetSampleIndex(wpc):  # WavPackUtils.cs:355
    return wpc.sample_index


def WavpackGetNumErrors(wpc):  # WavPackUtils.cs:363
    return wpc.crc_errors


def WavpackLossy(wpc):  # WavPackUtils.cs:371-374
    return bool(wpc.info.lossy_blocks) or (wpc.info.config_flags & CONFIG_HYBRID_FLAG) != 0


def WavpackGetSampleRate(wpc):  # WavPackUtils.cs:379-385
    i = wpc.info
    if i.sample_rate != 0:
        return int(i.dsd_multiplier) * int(i.sample_rate) * 8 if i.dsd_multiplier > 0 else int(i.sample_rate)
    return 44100


def WavpackGetNumChannels(wpc):  # WavPackUtils.cs:392-398
    return wpc.info.num_channels if wpc.info.num_channels != 0 else 2


def WavpackGetBitsPerSample(wpc):  # WavPackUtils.cs:409-415
    i = wpc.info
    if i.bits_per_sample != 0:
        return i.bits_per_sample // 8 if i.dsd_multiplier > 0 else i.bits_per_sample
    return 16


def WavpackGetBytesPerSample(wpc):  # WavPackUtils.cs:423-429
    return wpc.info.bytes_per_sample if wpc.info.bytes_per_sample != 0 else 2


def WavpackGetReducedChannels(wpc):  # WavPackUtils.cs:437-445
    i = wpc.info
    if i.reduced_channels != 0:
        return i.reduced_channels
    return i.num_channels if i.num_channels != 0 else 2


def WavpackGetFileFormat(wpc):  # WavPackUtils.cs:452
    return wpc.info.file_format


def WavpackGetFileExtension(wpc):  # WavPackUtils.cs:463-469
    e = wpc.info.file_extension.decode("utf-8", "replace")
    return e if e else "wav"


def WavpackGetErrorMessage(wpc):  # WavPackUtils.cs:471
    return wpc.error_message


def _stored(wpc, off, ln):
    return None if ln < 0 else wpc.data[off:off + ln].tobytes()


def WavpackGetHeader(wpc):  # WavPackUtils.cs:476
    return _stored(wpc, wpc.info.header_off, wpc.info.header_len)


def WavpackGetTrailer(wpc):  # WavPackUtils.cs:481
    return _stored(wpc, wpc.info.trailer_off, wpc.info.trailer_len)


def WavpackGetIsFive(wpc):  # WavPackUtils.cs:486
    return bool(wpc.info.five)


def WavpackGetVersion(wpc):  # WavPackUtils.cs:491
    return wpc.info.version


def WavpackGetIsFloat(wpc):  # WavPackUtils.cs:496
    return (wpc.info.config_flags & CONFIG_FLOAT_DATA) > 0


def WavpackGetMode(wpc):  # WavPackUtils.cs:133-167
    f = wpc.info.config_flags
    mode = 0
    if f & CONFIG_HYBRID_FLAG:
        mode |= MODE_HYBRID
    elif not (f & CONFIG_LOSSY_MODE):
        mode |= MODE_LOSSLESS
    if wpc.info.lossy_blocks:
        mode &= ~MODE_LOSSLESS
    if f & CONFIG_FLOAT_DATA:
        mode |= MODE_FLOAT
    if f & CONFIG_HIGH_FLAG:
        mode |= MODE_HIGH
        if (f & CONFIG_VERY_HIGH_FLAG) or wpc.info.version < 0x405:
            mode |= MODE_VERY_HIGH
    if f & CONFIG_FAST_FLAG:
        mode |= MODE_FAST
    if f & CONFIG_EXTRA_MODE:
        mode |= MODE_EXTRA | ((wpc.info.xmode << 12) & MODE_XMODE)
    if wpc.info.dsd_multiplier > 0:
        mode |= MODE_DSD
    return mode


def WavpackGetCompressionLevel(wpc):  # WavPackUtils.cs:169-187
    mode = WavpackGetMode(wpc)
    result = None
    if mode & MODE_FAST:
        result = "Fast"
    elif mode & MODE_VERY_HIGH:
        result = "Very High"
    elif mode & MODE_HIGH:
        result = "High"
    if mode & MODE_EXTRA:
        result = (result or "Default") + ", " + "Extra-%d" % ((mode & MODE_XMODE) >> 12)
    return result


def SetSample(wpc, sample):
    """WavPackUtils.cs:509-594 replaced by an O(1) move over the block index: the whole file is decoded
    once, so seeking is repositioning.  Returns False past the end, like the reference."""
    if wpc.info.total_samples >= 0 and sample >= wpc.info.total_samples:
        return False
    wpc.sample_index = max(0, int(sample))
    return True


def SetTime(wpc, milliseconds):  # WavPackUtils.cs:504-507
    return SetSample(wpc, milliseconds // 1000 * int(wpc.info.sample_rate))
