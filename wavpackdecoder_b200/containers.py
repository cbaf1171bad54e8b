"""Container headers the reference does not synthesise.

WvDemo writes a 44-byte RIFF/WAVE header when a file stores none (WvDemo.cs:78-105; `wvdemo.wave_header`), whatever
WavpackGetFileFormat says.  A file whose source was Sony Wave64 or Philips DSDIFF then comes out as a .wav with an
extension that says otherwise (WavpackGetFileExtension).  `wvdemo.unpack_files(..., container="native")` writes the
header of the format the file names instead, for the two formats whose audio layout is what the decoder already
produces (interleaved little-endian PCM; interleaved MSB-first DSD bytes):

  W64  'riff' / 'wave' / 'fmt ' / 'data' chunks identified by 16-byte GUIDs, 64-bit sizes that include the 24-byte chunk
       header, chunks padded to 8 bytes (Sony Wave64 specification).
  DFF  'FRM8' / 'DSD ' form with FVER, PROP('SND ': FS, CHNL, CMPR 'DSD ') and the 'DSD ' sound data chunk, big-endian
       64-bit sizes (DSDIFF 1.5 specification); sample data are the raw DSD bytes (WVB_OUT_DSD_RAW).

  CAF  'caff' file header, 'desc' (sample rate as a 64-bit float, 'lpcm', little-endian flag, bytes per packet, channels,
       bits) and 'data' (edit count 0) chunks with big-endian fields (Apple Core Audio Format specification); the PCM
       itself stays little-endian, which the format flags allow.

  DSF  'DSD ' (28 bytes), 'fmt ' (52 bytes: format version 1, DSD raw, channel type and count, one-bit sampling frequency,
       1 bit per sample, one-bit sample count per channel, 4096-byte blocks) and 'data' chunks, little-endian (Sony DSF
       specification 1.01).  DSF stores each channel in 4096-byte blocks, oldest bit in the LSB: the decoded bytes are re-laid
       out on the device (wvb_batch_dsd_to_dsf, wvb_dsf.cuh).

Like the rest of wvdemo.py this is plumbing: no sample is touched by host arithmetic.
"""
import struct

WP_FORMAT_WAV, WP_FORMAT_W64, WP_FORMAT_CAF, WP_FORMAT_DFF, WP_FORMAT_DSF = 0, 1, 2, 3, 4  # Defines.cs:148-155

_W64_TAIL = bytes.fromhex("f3acd3118cd100c04f8edb8a")
W64_RIFF = b"riff" + bytes.fromhex("2e91cf11a5d628db04c10000")
W64_WAVE = b"wave" + _W64_TAIL
W64_FMT = b"fmt " + _W64_TAIL
W64_DATA = b"data" + _W64_TAIL


def wave_format(num_channels, sample_rate, bits, byteps, is_float=False):
    """The 16-byte WAVEFORMAT body both RIFF and W64 carry (WaveHeader.cs field order)."""
    block_align = byteps * num_channels
    return struct.pack("<HHIIHH", 3 if is_float else 1, num_channels & 0xffff, sample_rate & 0xffffffff,
                       (sample_rate * block_align) & 0xffffffff, block_align & 0xffff, bits & 0xffff)


def w64_header(total_samples, num_channels, sample_rate, bits, byteps, is_float=False):
    """Sony Wave64 header for `total_samples` frames: 16+8 riff, 16 wave, 16+8+16 fmt, 16+8 data = 104 bytes."""
    data_bytes = total_samples * byteps * num_channels
    fmt = wave_format(num_channels, sample_rate, bits, byteps, is_float)
    fmt_chunk = W64_FMT + struct.pack("<Q", 24 + len(fmt)) + fmt  # 40 bytes: already a multiple of 8
    data_chunk_size = 24 + data_bytes
    total = 24 + 16 + len(fmt_chunk) + ((data_chunk_size + 7) & ~7)
    return W64_RIFF + struct.pack("<Q", total) + W64_WAVE + fmt_chunk + W64_DATA + struct.pack("<Q", data_chunk_size)


def w64_trailer(total_samples, num_channels, byteps):
    """Padding that brings the data chunk to a multiple of 8 bytes."""
    return bytes(-(total_samples * byteps * num_channels) % 8)


_DFF_CHANNEL_IDS = {1: [b"C000"], 2: [b"SLFT", b"SRGT"], 5: [b"MLFT", b"MRGT", b"C   ", b"LS  ", b"RS  "],
                    6: [b"MLFT", b"MRGT", b"C   ", b"LFE ", b"LS  ", b"RS  "]}


def dff_header(total_byte_times, num_channels, sample_rate):
    """DSDIFF header for `total_byte_times` bytes per channel; sample_rate is the one-bit rate per channel in Hz
    (what WavpackGetSampleRate reports for a DSD file, WavPackUtils.cs:379-385: 2 822 400 for DSD64)."""
    ids = _DFF_CHANNEL_IDS.get(num_channels) or [b"C%03d" % i for i in range(num_channels)]
    fver = b"FVER" + struct.pack(">QI", 4, 0x01050000)
    fs = b"FS  " + struct.pack(">QI", 4, sample_rate & 0xffffffff)
    chnl = b"CHNL" + struct.pack(">QH", 2 + 4 * num_channels, num_channels) + b"".join(ids)
    name = b"not compressed"
    cmpr = b"CMPR" + struct.pack(">Q", 4 + 1 + len(name) + 1) + b"DSD " + bytes([len(name)]) + name + b"\0"  # count + text padded to even
    prop_body = b"SND " + fs + chnl + cmpr
    prop = b"PROP" + struct.pack(">Q", len(prop_body)) + prop_body
    data_bytes = total_byte_times * num_channels
    form_size = 4 + len(fver) + len(prop) + 12 + data_bytes + (data_bytes & 1)
    return b"FRM8" + struct.pack(">Q", form_size) + b"DSD " + fver + prop + b"DSD " + struct.pack(">Q", data_bytes)


def dff_trailer(total_byte_times, num_channels):
    """IFF chunks are padded to an even length."""
    return bytes((total_byte_times * num_channels) & 1)


def caf_header(total_samples, num_channels, sample_rate, bits, byteps, is_float=False):
    """Core Audio Format header for `total_samples` frames of little-endian PCM: 8 + (12 + 32) + (12 + 4) = 68 bytes."""
    flags = 2 | (1 if is_float else 0)  # kCAFLinearPCMFormatFlagIsLittleEndian | ...IsFloat
    desc = struct.pack(">d4sIIIII", float(sample_rate), b"lpcm", flags, byteps * num_channels, 1, num_channels, bits)
    data_bytes = total_samples * byteps * num_channels
    return (b"caff" + struct.pack(">HH", 1, 0) + b"desc" + struct.pack(">q", len(desc)) + desc +
            b"data" + struct.pack(">qI", 4 + data_bytes, 0))


DSF_BLOCK = 4096
_DSF_CHANNEL_TYPE = {1: 1, 2: 2, 3: 3, 4: 4, 5: 6, 6: 7}  # mono, stereo, 3 channels, quad, 5 channels, 5.1 (4 channels: quad)


def dsf_data_bytes(total_byte_times, num_channels):
    return (total_byte_times + DSF_BLOCK - 1) // DSF_BLOCK * DSF_BLOCK * num_channels


def dsf_header(total_byte_times, num_channels, sample_rate):
    """Sony DSF header (92 bytes) for `total_byte_times` bytes per channel; sample_rate is the one-bit rate in Hz."""
    data = dsf_data_bytes(total_byte_times, num_channels)
    fmt = struct.pack("<4sQIIIIIIQII", b"fmt ", 52, 1, 0, _DSF_CHANNEL_TYPE.get(num_channels, 7), num_channels, sample_rate & 0xffffffff, 1,
                      total_byte_times * 8, DSF_BLOCK, 0)
    return struct.pack("<4sQQQ", b"DSD ", 28, 28 + 52 + 12 + data, 0) + fmt + struct.pack("<4sQ", b"data", 12 + data)
