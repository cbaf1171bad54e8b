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
  * WavpackUnpackSamples decodes a WINDOW of blocks on the device -- every block from the one that contains the current
    position on, all in parallel -- and later calls copy out of it; a seek (SetSample / SetTime) or a call with a different
    size starts a new window at the containing block, with the call grid the reference would have used (the grid decides
    mute granularity and short-weight casts on damaged streams);
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
        self.error_message = None
        self.crc_errors = 0
        self.sample_index = 0
        self.open_flags = 0
        self.device = 0
        self.lookahead_blocks = 0   # blocks decoded per device pass from the current position; 0 = to the end of the file
        self.decode_passes = 0      # device passes made so far (a streaming caller sees 1; every seek or chunk-size change adds one)
        self._out_channels = 0
        self._win = None            # the decoded window (_Table with .decoded)
        self._pending = None        # block table made by SetSample / SetTime, decoded by the next WavpackUnpackSamples
        self._dec = None

    def __del__(self):
        try:
            if self._dec is not None:
                self._dec.close()
        except Exception:
            pass


class _Table:
    """Block table of a window: the block decoding (re)starts at and the ones after it."""
    __slots__ = ("descs", "nblocks", "first", "landed", "nsamples", "chunk", "origin", "starts", "ends", "hdr", "decoded", "crc_flags",
                 "next_pos", "stopped_early", "lossy_blocks")


def WavpackOpenFileInput(infile, flags=0, device=0):
    """WavPackUtils.cs:36-120.  `infile`: bytes-like with the .wv stream.  Never raises for format errors: check
    WavpackGetErrorMessage(), like the reference.  Host work only (the block-header walk); nothing is decoded yet."""
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


def _make_table(wpc, chunk, origin, target=0, max_blocks=0):
    """Index pass for one window.  origin: ("open",) -- the table WavpackOpenFileInput + sequential reads produce;
    ("seek", wvb_seek_state or None) -- after the reference's seek() to `target`; ("regrid", previous call size) -- the caller
    went on reading at `target` with another call size (no seek: the block's consumed head was cut in the previous size)."""
    lib = N.load()
    d, size = wpc.data.ctypes.data, wpc.data.size
    info = N.FileInfo()
    n = C.c_size_t()
    first, landed = C.c_int64(0), C.c_int64(0)

    def call(descs, cap):
        if origin[0] == "open":
            rc = lib.wvb_index(d, size, wpc.open_flags, chunk, C.byref(info), descs, cap, C.byref(n))
            if rc == N.E_CAPACITY and max_blocks:
                rc = N.OK  # only the head of the table was asked for
            return rc
        state = C.byref(origin[1]) if origin[0] == "seek" and origin[1] is not None else None
        skip_chunk = origin[1] if origin[0] == "regrid" else 0
        return lib.wvb_index_seek(d, size, wpc.open_flags, state, target, skip_chunk, chunk, max_blocks, C.byref(info), descs, cap,
                                  C.byref(n), C.byref(first), C.byref(landed))

    if origin[0] == "open" and max_blocks:
        n.value = max_blocks
    elif call(None, 0) != N.OK:
        raise RuntimeError("block index failed")
    t = _Table()
    t.nblocks = n.value
    t.descs = (N.BlockDesc * max(t.nblocks, 1))()
    if t.nblocks and call(t.descs, t.nblocks) != N.OK:
        raise RuntimeError("block index failed")
    t.nblocks = min(t.nblocks, n.value)
    t.first, t.landed = first.value, landed.value
    t.chunk, t.origin = chunk, origin
    t.nsamples = int(info.indexed_samples) if t.nblocks else 0
    t.stopped_early, t.lossy_blocks = bool(info.stopped_early), info.lossy_blocks
    # per block: first / last+1 sample in the file, and what the reference's reader remembers of its header
    t.starts = np.array([t.first + int(t.descs[i].out_offset) for i in range(t.nblocks)], dtype=np.int64)
    t.ends = np.array([t.first + int(t.descs[i].out_offset) + int(t.descs[i].block_samples) for i in range(t.nblocks)], dtype=np.int64)
    t.hdr = [(int(b.in_offset), int(b.block_index), int(b.block_samples), int(b.in_bytes), int(b.avg_block_size)) for b in t.descs[:t.nblocks]]
    if origin[0] == "open" and max_blocks and t.nblocks:
        t.nsamples = int(t.ends[-1] - t.first)  # only the head of the whole-file table was kept
    t.decoded = None
    t.next_pos = t.landed if origin[0] != "open" else 0
    return t


def _reader_state(wpc):
    """What the reference's reader holds when seek() starts: the header it read last (the block of the last sample handed
    out, or the block decoding (re)started at) and the file position behind that block."""
    tab = wpc._pending or wpc._win
    if tab is None:
        tab = _make_table(wpc, SAMPLE_BUFFER_SIZE, ("open",), max_blocks=1)
    if tab.nblocks == 0:
        return None
    before = np.nonzero(tab.starts < wpc.sample_index)[0]
    j = int(before[-1]) if before.size else 0
    pos, block_index, block_samples, in_bytes, avg = tab.hdr[j]
    return N.SeekState(hdr_pos=pos, block_index=block_index, avg_block_size=avg, file_pos=pos + in_bytes, block_samples=block_samples,
                       ck_size=max(in_bytes - 8, 0))


def _decode_table(wpc, t):
    """One device pass over a window's blocks, all in parallel."""
    lib = N.load()
    from .batch import BatchDecoder
    nch = wpc._out_channels
    if t.nblocks:
        lib.wvb_rebase(t.descs, t.nblocks, 0, 0, N.OUT_INT32, 0)
        out = np.zeros(t.nsamples * nch + 16, dtype=np.int32)
        results = (N.BlockResult * t.nblocks)()
        slab = np.concatenate([wpc.data, np.zeros(64, dtype=np.uint8)])
        if wpc._dec is None:
            wpc._dec = BatchDecoder(wpc.device)  # raises without a CUDA device: no fallback
        wpc._dec.decode(slab.ctypes.data, slab.size, t.descs, t.nblocks, out.ctypes.data, t.nsamples * nch * 4, N.OUT_INT32, 0, results)
        wpc.decode_passes += 1
        t.decoded = out[: t.nsamples * nch]
        t.crc_flags = np.array([bool(results[i].rflags & N.RF_CRC_ERROR) for i in range(t.nblocks)], dtype=bool)
        wpc.info.lossy_blocks |= t.lossy_blocks
    else:
        t.decoded = np.zeros(0, dtype=np.int32)
        t.crc_flags = np.zeros(0, dtype=bool)
    wpc._win = t
    wpc._pending = None


def _new_window(wpc, chunk):
    win, pend = wpc._win, wpc._pending
    if pend is not None:  # after SetSample / SetTime
        t = pend if pend.chunk == chunk and pend.origin[0] == "seek" and not wpc.lookahead_blocks else \
            _make_table(wpc, chunk, pend.origin, pend.origin[2] if len(pend.origin) > 2 else wpc.sample_index, wpc.lookahead_blocks)
    elif win is None and wpc.sample_index == 0:
        t = _make_table(wpc, chunk, ("open",), max_blocks=wpc.lookahead_blocks)
    elif wpc.info.total_samples < 0:
        # unknown length: the reference cannot seek such files (WavPackUtils.cs:527) and neither can the index; a call-size
        # change mid-stream re-reads from the start with the new size
        t = _make_table(wpc, chunk, ("open",))
    else:
        prev = win.chunk if win is not None and win.next_pos == wpc.sample_index else 0
        t = _make_table(wpc, chunk, ("regrid", prev) if prev else ("seek", None), wpc.sample_index, wpc.lookahead_blocks)
    # a seek that starts by decoding an earlier block (see wvb_index_seek) counts that block's CRC verdict when the skip loop
    # finishes it (WavPackUtils.cs:273-275 inside the skip calls)
    _decode_table(wpc, t)
    if t.origin[0] == "seek" and t.nblocks:
        wpc.crc_errors += int(np.count_nonzero(t.crc_flags & (t.ends <= wpc.sample_index)))
    t.next_pos = wpc.sample_index


def WavpackUnpackSamples(wpc, buffer, samples):
    """WavPackUtils.cs:200-282: fills `buffer` (numpy int32, >= samples * channels) with right-justified samples and
    returns the number of complete samples unpacked (short at end of stream).

    The first call decodes every block from the current position on in ONE device pass (the whole file for a caller that
    starts at 0; `wpc.lookahead_blocks` bounds it) and later calls copy out of that window.  The call size is part of the
    reference's behaviour on damaged streams (mute granularity, short-weight casts): a call with another size, or after
    SetSample / SetTime, decodes a new window from the block that contains the current position.  With a bounded
    lookahead the call grid restarts at every window (exact for undamaged streams)."""
    if wpc.error_message:
        return 0
    samples = int(samples)
    if samples <= 0:
        return 0
    nch = wpc._out_channels
    win = wpc._win
    if win is None or wpc._pending is not None or win.chunk != samples or win.next_pos != wpc.sample_index:
        _new_window(wpc, samples)
        win = wpc._win
    done = 0
    while done < samples:
        n = int(min(samples - done, win.first + win.nsamples - wpc.sample_index))
        if wpc.info.total_samples >= 0 and wpc.sample_index < wpc.info.total_samples:
            n = int(min(n, wpc.info.total_samples - wpc.sample_index))  # the call returns at total_samples (WavPackUtils.cs:277)
        if n <= 0:
            if wpc.lookahead_blocks and win.nblocks >= wpc.lookahead_blocks:  # a full window: the stream may go on
                _new_window(wpc, samples)
                win = wpc._win
                if win.nsamples and wpc.sample_index < win.first + win.nsamples:
                    continue
            break
        a = (wpc.sample_index - win.first) * nch
        buffer[done * nch: (done + n) * nch] = win.decoded[a: a + n * nch]
        new_index = wpc.sample_index + n
        # crc_errors becomes visible when the block's last sample has been handed out (WavPackUtils.cs:273-275)
        wpc.crc_errors += int(np.count_nonzero(win.crc_flags & (win.ends > wpc.sample_index) & (win.ends <= new_index)))
        wpc.sample_index = new_index
        win.next_pos = new_index
        done += n
        if new_index == wpc.info.total_samples:
            break
    return done


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
    """WavPackUtils.cs:509-594.  The reference probes the file for the block that contains `sample` (up to 25 header reads
    steered by the average block size), restarts its decoder there and decodes-and-discards up to the target.  Here the
    index pass replays that probe sequence on the headers (wvb_index_seek; host work, no decoding) and the next
    WavpackUnpackSamples decodes from the block it ends on, the discarded head included, with the same call grid -- one
    device pass over the blocks from there on, not over the whole file.
    Returns False where the reference does: past the end of the stream, streams of unknown length, failed probes."""
    sample = int(sample)
    if sample >= wpc.info.total_samples or wpc.error_message:
        return False
    t = _make_table(wpc, wpc._win.chunk if wpc._win is not None else SAMPLE_BUFFER_SIZE, ("seek", _reader_state(wpc), sample), sample,
                    wpc.lookahead_blocks)
    if t.nblocks == 0:
        return False
    wpc._pending = t
    wpc.sample_index = t.landed
    return True


def SetTime(wpc, milliseconds):  # WavPackUtils.cs:504-507
    return SetSample(wpc, int(milliseconds) // 1000 * int(wpc.info.sample_rate))
