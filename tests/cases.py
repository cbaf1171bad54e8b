"""Shared synthetic-stream cases (name, open_flags, chunk, encoder kwargs) used by the CPU (emulation)
and GPU parity tests.  Edge cases follow the reference's own branches: mono / FALSE_STEREO, every
decorrelation term class, 8/16/24/32-bit, shift, INT32 with and without WVX, float flag, hybrid
(bitrate, balance), 5.1 via OPEN_2CH_MAX, odd chunk sizes, tiny blocks, optional sub-blocks."""
from _harness import KIND_FLOAT, KIND_HYBRID

T16 = [18, 18, 2, 3, -2, 18, 2, 4, 7, 5, 3, 6, 8, -1, 18, 2]

PCM_CASES = [
    ("default", 0, 4096, dict()),
    ("mono", 0, 4096, dict(channels=1)),
    ("false_stereo", 0, 4096, dict(false_stereo=1, seconds=1.5)),
    ("no_joint", 0, 4096, dict(joint_stereo=0)),
    ("terms_all_pos", 0, 4096, dict(terms=[1, 2, 3, 4, 5, 6, 7, 8, 17, 18])),
    ("terms_neg", 0, 4096, dict(terms=[-1, -2, -3, 18, -1, 2])),
    ("terms16", 0, 4096, dict(terms=T16, deltas=[2] * 16)),
    # state above 128 words per thread: the software-pipelined pass loop and the one-warp CTAs (every term class in it)
    ("terms16_all_classes", 0, 4096, dict(terms=[1, 2, 3, 4, 5, 6, 7, 8, 17, 18, -1, -2, -3, 8, 7, 6], deltas=[1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3, 4, 5, 6, 7, 2])),
    ("terms16_mono", 0, 4096, dict(channels=1, terms=[17, 18, 1, 8, 5, 7, 6, 8, 8, 7, 6, 5, 8, 8, 8, 8], deltas=[2] * 16)),
    ("terms15_odd_count_24bit", 0, 1000, dict(bits=24, terms=[18, 8, 7, 6, 5, 8, 7, 6, -1, 8, 17, 8, -3, 7, 8], deltas=[3] * 15, seconds=0.5)),
    ("terms16_hybrid", 0, 4096, dict(kind=KIND_HYBRID, terms=T16, deltas=[2] * 16, seconds=0.5)),
    ("terms16_int32_wvx", 0, 4096, dict(bits=32, int32_sent_bits=8, terms=T16, deltas=[2] * 16, seconds=0.5)),
    ("terms0", 0, 4096, dict(terms=[])),
    ("delta7", 0, 4096, dict(terms=[18, 3, -3], deltas=[7, 0, 5])),
    ("mono_terms", 0, 4096, dict(channels=1, terms=[17, 18, 1, 8, 5])),
    ("8bit", 0, 4096, dict(bits=8)),
    ("8bit_mono", 0, 4096, dict(bits=8, channels=1)),
    ("24bit", 0, 4096, dict(bits=24, sample_rate=48000, block_samples=24000)),
    ("24bit_mono", 0, 4096, dict(bits=24, channels=1)),
    ("shift4_16", 0, 4096, dict(bits=16, shift=4)),
    ("small_blocks", 0, 4096, dict(block_samples=100, seconds=0.5)),
    ("ragged_tail", 0, 4096, dict(nsamples=50001, block_samples=7001)),
    ("one_sample", 0, 4096, dict(nsamples=1)),
    ("fifteen_samples", 0, 4096, dict(nsamples=15)),
    ("chunk1000", 0, 1000, dict()),
    ("chunk7", 0, 7, dict(seconds=0.1)),
    ("chunk17", 0, 17, dict(seconds=0.1, channels=1)),
    ("int32_wvx", 0, 4096, dict(bits=32, int32_sent_bits=8)),
    ("int32_wvx_new", 0, 4096, dict(bits=32, int32_sent_bits=8, int32_new_wvx=1, int32_max_width=0)),
    ("int32_wvx_new_mw", 0, 4096, dict(bits=32, int32_sent_bits=8, int32_new_wvx=1, int32_max_width=28)),
    ("int32_nowvx", 0, 4096, dict(bits=32, int32_sent_bits=8, int32_wvx=0)),
    ("int32_zeros", 0, 4096, dict(bits=32, int32_sent_bits=6, int32_zeros=3)),
    ("int32_ones", 0, 4096, dict(bits=32, int32_sent_bits=6, int32_ones=2)),
    ("int32_dups", 0, 4096, dict(bits=32, int32_sent_bits=6, int32_dups=2)),
    ("int32_only_zeros", 0, 4096, dict(bits=32, int32_sent_bits=0, int32_zeros=9)),
    ("int32_mono", 0, 4096, dict(bits=32, int32_sent_bits=8, channels=1)),
    ("hybrid", 0, 4096, dict(kind=KIND_HYBRID, terms=[18, 18, 2, 3])),
    ("hybrid_mono", 0, 4096, dict(kind=KIND_HYBRID, channels=1, terms=[18, 17, 2])),
    ("hybrid_neg", 0, 4096, dict(kind=KIND_HYBRID)),
    # the stock term lists under hybrid / float / int32: fixup and entropy code over the in-register decorrelators
    ("hybrid_mono_stock_terms", 0, 4096, dict(kind=KIND_HYBRID, channels=1, terms=[18, 18, 2, 3], deltas=[2] * 4)),
    ("hybrid_stock_terms_balance_chunk999", 0, 999, dict(kind=KIND_HYBRID, hybrid_balance=1)),
    ("hybrid_stock_terms_v402_24bit", 0, 4096, dict(kind=KIND_HYBRID, version=0x402, bits=24, hybrid_bitrate=6 * 256)),
    ("float_mono_stock_terms", 0, 4096, dict(kind=KIND_FLOAT, bits=32, channels=1, terms=[18, 18, 2, 3], deltas=[2] * 4)),
    ("int32_mono_stock_terms", 0, 4096, dict(bits=32, int32_sent_bits=8, channels=1, terms=[18, 18, 2, 3], deltas=[2] * 4)),
    ("hybrid_balance", 0, 4096, dict(kind=KIND_HYBRID, hybrid_balance=1, terms=[18, 2])),
    ("hybrid_2bit", 0, 4096, dict(kind=KIND_HYBRID, hybrid_bitrate=512, terms=[18, 2])),
    ("hybrid_8bit_src", 0, 4096, dict(kind=KIND_HYBRID, bits=8, hybrid_bitrate=768, terms=[17])),
    ("float", 0, 4096, dict(kind=KIND_FLOAT, bits=32)),
    ("float_shift", 0, 4096, dict(kind=KIND_FLOAT, bits=32, float_shift=3, float_max_exp=126)),
    ("float_shiftneg", 0, 4096, dict(kind=KIND_FLOAT, bits=32, float_max_exp=120, float_new_wvx=1)),
    ("float_lossyflag", 0, 4096, dict(kind=KIND_FLOAT, bits=32, float_flags=0x20)),
    ("5.1_2ch", 8, 4096, dict(channels=6, bits=24, sample_rate=48000, block_samples=24000, terms=T16, deltas=[2] * 16, seconds=1.2)),
    ("extras", 0, 4096, dict(extras=1 | 2 | 4 | 8 | 16 | 32 | 64, sample_rate=37800)),
    ("unknown_len", 0, 4096, dict(unknown_length=1)),
    ("all_hist_same_terms", 0, 4096, dict(extras=128, terms=[2, 2, 2])),
    ("false_stereo_24bit", 0, 4096, dict(false_stereo=1, bits=24, seconds=1.2)),
    ("false_stereo_int32_nowvx", 0, 4096, dict(false_stereo=1, bits=32, int32_sent_bits=8, int32_wvx=0, seconds=1.2)),
    ("false_stereo_float", 0, 4096, dict(false_stereo=1, kind=KIND_FLOAT, bits=32, float_shift=2, seconds=1.2)),
    ("false_stereo_hybrid", 0, 4096, dict(false_stereo=1, kind=KIND_HYBRID, terms=[18, 2], seconds=1.2)),
    ("v402_hybrid_history_skip", 0, 4096, dict(kind=KIND_HYBRID, version=0x402, terms=[18, 2])),
    ("v402_hybrid_mono", 0, 4096, dict(kind=KIND_HYBRID, version=0x402, channels=1, terms=[3])),
    ("v407", 0, 4096, dict(version=0x407)),
    ("delta0", 0, 4096, dict(terms=[18, 18, 2, 3, -2], deltas=[0, 0, 0, 0, 0])),
    ("default_terms_delta5", 0, 4096, dict(terms=[18, 18, 2, 3, -2], deltas=[5, 4, 3, 2, 1])),
    ("mono_default_terms", 0, 4096, dict(channels=1, terms=[18, 18, 2, 3])),
    ("terms_minus3_only", 0, 4096, dict(terms=[-3])),
    ("shift8_24bit", 0, 4096, dict(bits=24, shift=8)),
    ("hybrid_24bit", 0, 4096, dict(kind=KIND_HYBRID, bits=24, hybrid_bitrate=6 * 256, terms=[18, 18, 2])),
    ("big_block", 0, 4096, dict(block_samples=88200, seconds=2.0)),
    # the short lists of small terms libwavpack/FFmpeg write in their fast and default modes
    ("ff_list_ff_default", 0, 4096, dict(terms=[18, 18, 2, 17, 3])),
    ("ff_list_ff_nojoint", 0, 4096, dict(terms=[18, 17, -1, 3, 2], joint_stereo=0)),
    ("ff_list_17_18_18_m2_2", 0, 4096, dict(terms=[17, 18, 18, -2, 2], deltas=[2, 3, 4, 5, 6])),
    ("ff_list_17_17_m2_2_3", 0, 4096, dict(terms=[17, 17, -2, 2, 3])),
    ("ff_list_fast", 0, 4096, dict(terms=[18, 17])),
    ("ff_list_all_cross", 0, 4096, dict(terms=[-3, -1, 1, -2, -3], deltas=[1, 2, 3, 4, 5])),
    ("ff_list_one_term", 0, 4096, dict(terms=[1])),
    ("ff_list_mono", 0, 4096, dict(channels=1, terms=[18, 17, 3, 2, 1])),
    ("ff_list_false_stereo", 0, 4096, dict(false_stereo=1, terms=[17, 18, 18, 1, 2], seconds=1.2)),
    # odd block lengths: block outputs start at byte offsets that are not multiples of four (packed output writer head/tail)
    ("8bit_mono_odd_blocks", 0, 4096, dict(bits=8, channels=1, block_samples=1001, seconds=0.25)),
    ("8bit_stereo_odd_blocks", 0, 4096, dict(bits=8, block_samples=1001, seconds=0.25)),
    ("16bit_mono_odd_blocks", 0, 4096, dict(channels=1, block_samples=1001, seconds=0.25)),
    ("24bit_mono_odd_blocks", 0, 4096, dict(bits=24, channels=1, block_samples=1001, seconds=0.25)),
    ("24bit_stereo_odd_blocks", 0, 4096, dict(bits=24, block_samples=1001, seconds=0.25)),
    ("24bit_mono_tiny_blocks", 0, 4096, dict(bits=24, channels=1, block_samples=1, seconds=0.002)),
    ("8bit_mono_tiny_blocks", 0, 3, dict(bits=8, channels=1, block_samples=3, seconds=0.003)),
    ("list_six_terms", 0, 4096, dict(terms=[18, 18, 2, 17, 3, 1])),
    ("list_term4", 0, 4096, dict(terms=[18, 4, 2])),
]

from _harness import KIND_DSD  # noqa: E402

DSD_CASES = [
    ("dsd0_mono", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, channels=1, seconds=0.1, block_samples=8192)),
    ("dsd0_stereo", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, seconds=0.1, block_samples=8192)),
    ("dsd1_mono", 0, 4096, dict(kind=KIND_DSD, dsd_mode=1, channels=1, seconds=0.1, block_samples=8192)),
    ("dsd1_stereo", 0, 4096, dict(kind=KIND_DSD, dsd_mode=1, seconds=0.1, block_samples=8192)),
    ("dsd1_raw_probs", 0, 4096, dict(kind=KIND_DSD, dsd_mode=1, dsd_raw_probs=1, dsd_history_bits=2, seconds=0.1, block_samples=8192)),
    ("dsd1_h0_chunk1000", 0, 1000, dict(kind=KIND_DSD, dsd_mode=1, dsd_history_bits=0, seconds=0.1, block_samples=5000)),
    ("dsd1_h5", 0, 4096, dict(kind=KIND_DSD, dsd_mode=1, dsd_history_bits=5, seconds=0.15, block_samples=22050)),
    ("dsd3_mono", 0, 4096, dict(kind=KIND_DSD, dsd_mode=3, channels=1, seconds=0.1, block_samples=8192)),
    ("dsd3_stereo", 0, 4096, dict(kind=KIND_DSD, dsd_mode=3, seconds=0.1, block_samples=8192)),
    ("dsd3_rate200", 0, 4096, dict(kind=KIND_DSD, dsd_mode=3, dsd_rate_i=200, seconds=0.1, block_samples=5000)),
    ("dsd3_extras", 0, 4096, dict(kind=KIND_DSD, dsd_mode=3, seconds=0.1, block_samples=8192, extras=2 | 4 | 8 | 16)),
    # raw-mode blocks whose payload and output start at every kind of misalignment, shorter than one vector step, and long
    ("dsd0_mono_odd_blocks", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, channels=1, seconds=0.1, block_samples=5001)),
    ("dsd0_stereo_odd_blocks", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, seconds=0.1, block_samples=4099, extras=2 | 4)),
    ("dsd0_mono_tiny_blocks", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, channels=1, seconds=0.002, block_samples=37)),
    ("dsd0_stereo_big_blocks", 0, 4096, dict(kind=KIND_DSD, dsd_mode=0, seconds=0.2, block_samples=44100)),
]


def _blocks(data):
    import struct
    off, out = 0, []
    while off + 32 <= len(data):
        cks = struct.unpack_from("<I", data, off + 4)[0]
        out.append((off, cks + 8))
        off += cks + 8
    return out


# damaged streams whose decode depends on state an EARLIER block left in the reference's decoder (quirk C-8): the device
# decodes every block from its own bytes, so these must come back FLAGGED (WVB_RF_INEXACT) and never fault; samples may differ
INEXACT_BY_DESIGN = {"dsd1_truncated", "dsd3_truncated", "block2_terms_to_dummy",
                     "block2_entropy_to_dummy", "mono_block1_terms_to_dummy", "block2_terms_dummy_first_block_too"}


def _subblocks(data, off):
    """(id byte offset, id, payload offset, padded payload length) of every sub-block of the block at `off`."""
    import struct
    cks = struct.unpack_from("<I", data, off + 4)[0]
    at, end, out = off + 32, off + 8 + cks, []
    while at + 2 <= end:
        idb, words, hdr = data[at], data[at + 1], 2
        if idb & 0x80:
            words |= (data[at + 2] << 8) | (data[at + 3] << 16)
            hdr = 4
        out.append((at, idb & 0x3f, at + hdr, words * 2))
        at += hdr + words * 2
    return out


def corrupt_cases():
    """(name, bytes, open_flags, chunk): damaged streams, the domain's "erasures".  Expected behaviour is whatever the
    oracle does: mute from the start of the caller chunk, CRC error counts, truncated output, gap zero fill, 0x55 DSD mute."""
    from _harness import make_file
    out = []
    cfg, src, data = make_file(seconds=2.0)
    bl = _blocks(data)

    def flip(d, at, mask=0x40):
        d = bytearray(d)
        d[at] ^= mask
        return bytes(d)

    mid1 = bl[1][0] + bl[1][1] // 2
    out.append(("flip_bitstream", flip(data, mid1), 0, 4096))
    out.append(("flip_bitstream_chunk1000", flip(data, mid1), 0, 1000))
    out.append(("flip_header_crc", flip(data, bl[2][0] + 28, 1), 0, 4096))
    out.append(("truncated_midblock", data[:bl[2][0] + bl[2][1] // 2], 0, 4096))
    out.append(("truncated_in_header", data[:bl[2][0] + 20], 0, 4096))
    out.append(("dropped_block_gap", data[:bl[1][0]] + data[bl[2][0]:], 0, 4096))
    out.append(("garbage_between_blocks", data[:bl[1][0]] + b"\x00wvpk garbage 12345" * 3 + data[bl[1][0]:], 0, 4096))
    out.append(("flip_tail_in_silence", flip(data, bl[1][0] + bl[1][1] - 3, 0xff), 0, 4096))
    d = bytearray(data)
    for k in range(200):
        d[bl[1][0] + 300 + k] = 0xff
    out.append(("ones_burst", bytes(d), 0, 4096))
    cfg, src, datam = make_file(seconds=2.0, channels=1)
    blm = _blocks(datam)
    out.append(("mono_flip", flip(datam, blm[1][0] + blm[1][1] // 2, 0x10), 0, 4096))
    out.append(("mono_flip_chunk777", flip(datam, blm[1][0] + blm[1][1] // 2, 0x10), 0, 777))
    # quirk C-5 (UnpackUtils.cs:572-575): a magnitude fault in the first sample of a mono block that begins 1000 samples
    # into the caller's call: buffer index 1000 + 0 == sample_count 1000, so the reference mutes nothing
    cfg, src, datac5 = make_file(channels=1, seconds=0.3, block_samples=1000, terms=[18, 18, 2, 3], deltas=[2] * 4)
    blc = _blocks(datac5)
    for byte, bit, chunk in ((47, 0, 4096), (49, 4, 4096), (47, 4, 4096), (49, 4, 2000)):
        out.append(("mono_fault_at_coincident_buffer_index_%d_%d_%d" % (byte, bit, chunk), flip(datac5, blc[1][0] + byte, 1 << bit), 0, chunk))
    cfg, src, datah = make_file(seconds=2.0, kind=1)
    blh = _blocks(datah)
    out.append(("hybrid_flip", flip(datah, blh[1][0] + blh[1][1] // 2, 0x10), 0, 4096))
    cfg, src, datai = make_file(seconds=2.0, bits=32, int32_sent_bits=8)
    bli = _blocks(datai)
    out.append(("int32_flip_main", flip(datai, bli[1][0] + 2000, 0x10), 0, 4096))
    for m in (0, 1, 3):
        cfg, src, dd = make_file(kind=KIND_DSD, dsd_mode=m, seconds=0.3, block_samples=20000)
        bd = _blocks(dd)
        out.append(("dsd%d_flip" % m, flip(dd, bd[1][0] + bd[1][1] // 2, 0x04), 0, 4096))
        out.append(("dsd%d_flip_chunk3000" % m, flip(dd, bd[1][0] + bd[1][1] // 2, 0x04), 0, 3000))
        # last block cut inside its ID_DSD_BLOCK payload: the reference keeps the previous block's (exhausted) DSD state
        out.append(("dsd%d_truncated" % m, dd[:bd[-1][0] + 600], 0, 4096))
    # metadata failure in the middle of a stream: ID_ENCODER_INFO (not optional) as first sub-block of block 1
    bad = bytearray(data)
    bad[bl[1][0] + 32] = 0x01
    out.append(("bad_metadata_id_midstream", bytes(bad), 0, 4096))
    out.append(("truncated_in_metadata", data[:bl[2][0] + 40], 0, 4096))
    # sub-block surgery: a block that loses one of its sub-blocks decodes with the state the previous block left behind.
    # weights without terms used to index shared memory at -1 on the device: the terms id becomes ID_DUMMY
    for name, sub_id in (("terms", 2), ("weights", 3), ("samples", 4), ("entropy", 5)):
        d = bytearray(data)
        at = [s for s in _subblocks(data, bl[2][0]) if s[1] == sub_id][0][0]
        d[at] &= ~0x3f
        out.append(("block2_%s_to_dummy" % name, bytes(d), 0, 4096))
    dm = bytearray(datam)
    dm[[s for s in _subblocks(datam, blm[1][0]) if s[1] == 2][0][0]] &= ~0x3f
    out.append(("mono_block1_terms_to_dummy", bytes(dm), 0, 4096))
    # a shorter term list with the longer weights sub-block of the original (count checked against the carried terms only when longer)
    d = bytearray(data)
    t_at, _, t_pay, t_len = [s for s in _subblocks(data, bl[2][0]) if s[1] == 2][0]
    d[t_at] = 0x00  # the 5-term list becomes a dummy ...
    out.append(("block2_terms_dummy_first_block_too", bytes(d[:bl[1][0]]) + bytes(d[bl[2][0]:]), 0, 4096))
    return out
