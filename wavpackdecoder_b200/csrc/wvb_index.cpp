// wvb_index.cpp -- host-side block index pass of libwvb (no CUDA in this file).
//
// Replaces, for the batch decoder, the reference's serial stream driver:
//   read_next_header            WavPackUtils.cs:600-671   (header scan)
//   unpack_init + metadata walk UnpackUtils.cs:24-68, MetadataUtils.cs:15-193
//   WavpackOpenFileInput        WavPackUtils.cs:36-120    (config derivation)
//   WavpackUnpackSamples        WavPackUtils.cs:200-282   (block sequencing, gaps, call/chunk grid)
// It never touches sample data: the payload of the decode-relevant sub-blocks is only
// located (offset/length) and validated the way the reference's readers validate it; the
// contents are parsed on the device.  Output: one wvb_block_desc per block the reference
// would decode, in output order.
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wvb.h"

namespace {

enum : uint32_t {
    F_BYTES_STORED = 3, F_MONO = 4, F_HYBRID = 8, F_JOINT = 0x10, F_FLOAT = 0x80, F_INT32 = 0x100, F_HYB_BITRATE = 0x200,
    F_HYB_BALANCE = 0x400, F_INITIAL = 0x800, F_FINAL = 0x1000, F_FALSE_STEREO = 0x40000000u, F_DSD = 0x80000000u
};
constexpr uint32_t SHIFT_LSB = 13, SRATE_LSB = 23;
constexpr uint32_t SHIFT_MASK = 0x1fu << SHIFT_LSB, SRATE_MASK = 0xfu << SRATE_LSB;
constexpr int READ_BUFFER = 16 * 1024; // Defines.cs:20 BITSTREAM_BUFFER_SIZE (shared read_buffer, WavpackContext.cs:17)

const int64_t kSampleRates[] = {6000, 8000, 9600, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 64000, 88200, 96000, 192000};

struct Header { // WavpackHeader.cs:15-22
    uint32_t ckSize = 0;
    int version = 0;
    int64_t total_samples = 0, block_index = 0;
    uint32_t block_samples = 0, flags = 0;
    int32_t crc = 0;
    size_t pos = 0; // stream_position
    int64_t avg = 0; // average_block_size (WavPackUtils.cs:647-650), folded over every header this reader accepted
};

struct Ctx { // the parts of WavpackContext/WavpackStream the index pass must carry between blocks
    const uint8_t *d;
    size_t len;
    size_t pos = 0; // infile position
    Header hdr;
    wvb_file_info *info;
    // persistent stream state (quirk C-8)
    int num_terms = 0;
    int8_t terms[16] = {0};
    bool terms_from_this_block = false;
    uint8_t int32_info[4] = {0, 0, 0, 0};
    uint8_t float_info[4] = {0, 0, 0, 0};
    bool have_int32 = false, have_float = false;
    bool wvx_present = false; // wps.wvxbits != null
    bool wvbits_nonempty = false; // wps.wvbits.end != 0
    bool dsd_ready = false;
    int64_t sample_index = 0;
    // current block's descriptor under construction
    wvb_block_desc cur;
    bool cur_inexact = false;
    bool exception = false; // the C# code would have thrown
};

inline bool valid_header(const uint8_t *b) // WavPackUtils.cs:632
{
    return b[0] == 'w' && b[1] == 'v' && b[2] == 'p' && b[3] == 'k' && (b[4] & 1) == 0 && b[6] < 16 && b[7] == 0 && b[9] == 4 &&
           b[8] >= 0x02 && b[8] <= 0x10;
}

inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// Forward scan for the next acceptable header; same acceptance test, same 1 MiB give-up rule.
bool read_next_header(Ctx &c)
{
    size_t p = c.pos;
    int64_t skipped = 0;
    for (;;) {
        if (p + 32 > c.len) { c.pos = c.len; return false; }
        const uint8_t *b = c.d + p;
        if (valid_header(b)) {
            Header &h = c.hdr;
            h.ckSize = le32(b + 4);
            h.version = b[8] | (b[9] << 8);
            h.total_samples = (int64_t)(((uint64_t)b[11] << 32) | le32(b + 12));
            h.block_index = (int64_t)(((uint64_t)b[10] << 32) | le32(b + 16));
            h.block_samples = le32(b + 20);
            h.flags = le32(b + 24);
            h.crc = (int32_t)le32(b + 28);
            h.pos = p;
            h.avg = h.avg == 0 ? (int64_t)h.ckSize : (h.avg + (int64_t)h.ckSize) / 2;
            c.pos = p + 32;
            return true;
        }
        size_t q = p + 1;
        while (q < p + 32 && c.d[q] != 'w') q++;
        skipped += (int64_t)(q - p);
        p = q;
        if (skipped > 1048576) { c.pos = p; return false; }
    }
}

int hist_entries_consumed(int term, bool stereo) // bytes one history entry occupies (UnpackUtils.cs:289-348)
{
    if (term > 8) return stereo ? 8 : 4;
    if (term < 0) return 4;
    return term * (stereo ? 4 : 2);
}

// DSD "fast" table validation, same acceptance rules as init_dsd_block_fast (DsdUtils.cs:149-242)
bool dsd_fast_valid(const uint8_t *p, size_t n, size_t at)
{
    if (at == n) return false;
    int history_bits = p[at++];
    if (at == n || history_bits > 5) return false;
    int bins = 1 << history_bits;
    size_t tot = (size_t)256 * bins;
    std::vector<uint8_t> prob(tot, 0);
    int max_probability = p[at++];
    if (max_probability < 0xFF) {
        size_t outp = 0;
        while (outp < tot && at < n) {
            int code = p[at++];
            if (code > max_probability) {
                int z = code - max_probability;
                while (outp < tot && z-- > 0) prob[outp++] = 0;
            } else if (code != 0)
                prob[outp++] = (uint8_t)code;
            else
                break;
        }
        if (outp < tot || (at < n && p[at++] > 0)) return false;
    } else if (n - at > tot) {
        memcpy(prob.data(), p + at, tot);
        at += tot;
    } else
        return false;
    int total = 0;
    for (int b = 0; b < bins; b++) {
        unsigned sum = 0;
        for (int i = 0; i < 256; i++) sum = (sum + prob[(size_t)b * 256 + i]) & 0xffff;
        if (sum) {
            total += (int)sum;
            if (total > bins * 1280) return false;
        }
    }
    if (n - at < 4 || total > bins * 1280) return false;
    return true;
}

// unpack_init (UnpackUtils.cs:24-68): walks the sub-blocks of the block whose header was just
// read, fills c.cur, updates the carried state.  Returns false where the reference returns FALSE.
bool unpack_init(Ctx &c)
{
    const Header &h = c.hdr;
    wvb_file_info &I = *c.info;
    wvb_block_desc &B = c.cur;
    memset(&B, 0, sizeof(B));
    B.in_offset = h.pos;
    B.in_bytes = (uint32_t)std::min<uint64_t>((uint64_t)h.ckSize + 8, c.len - h.pos); // a truncated last block is still decoded (muted)
    B.block_samples = h.block_samples;
    B.flags = h.flags;
    B.crc = h.crc;
    B.block_index = h.block_index;
    B.version = (uint16_t)h.version;
    c.cur_inexact = false;
    c.terms_from_this_block = false;

    if (h.block_samples > 0 && h.block_index != 0xFFFFFFFFLL) c.sample_index = h.block_index;

    const bool stereo = (h.flags & (F_MONO | F_FALSE_STEREO)) == 0;
    int64_t bytecount = 24;
    bool hasdata = false;
    bool terms_seen = false, weights_ok = true;
    auto fail_id = [&](int id) {
        snprintf(I.error_message, sizeof(I.error_message), "invalid metadata id %d", id);
        return false;
    };

    while (bytecount < (int64_t)h.ckSize) {
        if (c.pos + 2 > c.len) { c.pos = c.len; break; } // ReadByte throws -> read_metadata_buff returns FALSE
        int id = c.d[c.pos], words = c.d[c.pos + 1];
        c.pos += 2;
        bytecount += 2;
        int64_t byte_length = (int64_t)words << 1;
        if (id & 0x80) {
            id &= ~0x80;
            if (c.pos + 2 > c.len) { c.pos = c.len; break; }
            byte_length += (int64_t)c.d[c.pos] << 9;
            byte_length += (int64_t)c.d[c.pos + 1] << 17;
            c.pos += 2;
            bytecount += 2;
        }
        int64_t bytes_to_read = byte_length;
        if (id & 0x40) { id &= ~0x40; byte_length--; }
        size_t data_at = c.pos;
        if (byte_length == 0)
            hasdata = false;
        else {
            bytecount += bytes_to_read;
            if (bytes_to_read > 0) {
                if (c.pos + (size_t)bytes_to_read > c.len) { c.pos = c.len; break; } // short Read -> FALSE
                c.pos += (size_t)bytes_to_read;
                hasdata = true;
            }
        }
        const uint8_t *p = c.d + data_at;
        const uint32_t rel = (uint32_t)(data_at - h.pos);
        // a C# byte[] .Length for copy_data'd payloads (WavpackMetadata.cs:25-36, quirk C-10)
        const int64_t array_len = bytes_to_read <= READ_BUFFER ? byte_length : bytes_to_read;
        const bool copyable = hasdata && byte_length > 0;

        switch (id) { // process_metadata, MetadataUtils.cs:111-193
        case 0x00: break; // ID_DUMMY
        case 0x02: { // read_decorr_terms, UnpackUtils.cs:156-187
            int termcnt = (int)byte_length;
            if (termcnt > 16) return fail_id(id);
            if (termcnt < 0) { c.exception = true; return false; }
            int8_t t[16];
            for (int k = 0; k < termcnt; k++) {
                int term = (int)(p[k] & 0x1f) - 5;
                t[termcnt - 1 - k] = (int8_t)term;
                if (term < -3 || (term > 8 && term < 17) || term > 18) return fail_id(id);
                if (term == 0) c.cur_inexact = true; // accepted by the reference, never emitted by an encoder; not reproduced on the device
            }
            memcpy(c.terms, t, sizeof(t));
            c.num_terms = termcnt;
            c.terms_from_this_block = true;
            terms_seen = true;
            B.sub_off[WVB_SUB_TERMS] = rel; B.sub_len[WVB_SUB_TERMS] = (uint32_t)termcnt;
            break;
        }
        case 0x03: { // read_decorr_weights, UnpackUtils.cs:196-239
            int termcnt = (int)byte_length;
            if (stereo) termcnt /= 2;
            if (termcnt > c.num_terms) return fail_id(id);
            if (!terms_seen) weights_ok = false; // weights applied to passes carried over from an earlier block
            B.sub_off[WVB_SUB_WEIGHTS] = rel; B.sub_len[WVB_SUB_WEIGHTS] = (uint32_t)std::max<int64_t>(byte_length, 0);
            break;
        }
        case 0x04: { // read_decorr_samples, UnpackUtils.cs:250-360
            // replay the parse loop's cursor to see whether it stays inside the payload and the pass array (quirk C-1)
            int64_t counter = 0;
            int term = c.num_terms > 0 ? c.terms[c.num_terms - 1] : 0;
            int dpp_index = c.num_terms - 1;
            if (h.version == 0x402 && (h.flags & F_HYBRID)) counter += stereo ? 4 : 2;
            const int64_t buf_len = bytes_to_read <= READ_BUFFER ? READ_BUFFER : bytes_to_read;
            while (counter < byte_length) {
                int step = hist_entries_consumed(term, stereo);
                if (counter + step > buf_len || dpp_index < 0) { c.exception = true; return false; }
                if (counter + step > byte_length) c.cur_inexact = true; // parsed stale bytes of the shared read buffer
                counter += step;
                dpp_index--;
                if (step == 0) { c.exception = true; return false; }
            }
            if (!terms_seen) weights_ok = false;
            B.sub_off[WVB_SUB_SAMPLES] = rel; B.sub_len[WVB_SUB_SAMPLES] = (uint32_t)std::max<int64_t>(byte_length, 0);
            break;
        }
        case 0x05: // read_entropy_vars, WordsUtils.cs:75-116
            if (byte_length != 12 && stereo) return fail_id(id);
            if (byte_length < 6) c.cur_inexact = true; // medians read from stale buffer bytes
            B.sub_off[WVB_SUB_ENTROPY] = rel; B.sub_len[WVB_SUB_ENTROPY] = (uint32_t)std::max<int64_t>(byte_length, 0);
            break;
        case 0x06: { // read_hybrid_profile, WordsUtils.cs:124-187
            int64_t k = 0;
            if (h.flags & F_HYB_BITRATE) k += stereo ? 4 : 2;
            k += stereo ? 4 : 2;
            if (k > byte_length) c.cur_inexact = true;
            if (k < byte_length) {
                k += stereo ? 4 : 2;
                if (k < byte_length) return fail_id(id);
                if (k > byte_length) c.cur_inexact = true;
            }
            if (!B.sub_off[WVB_SUB_ENTROPY]) c.cur_inexact = true; // profile written into a words_data carried from an earlier block
            B.sub_off[WVB_SUB_HYBRID] = rel; B.sub_len[WVB_SUB_HYBRID] = (uint32_t)std::max<int64_t>(byte_length, 0);
            break;
        }
        case 0x07: break; // ID_SHAPING_WEIGHTS ignored
        case 0x08: // read_float_info, FloatUtils.cs:15-30
            if (byte_length != 4) return fail_id(id);
            memcpy(c.float_info, p, 4);
            c.have_float = true;
            I.float_norm_exp = p[3];
            break;
        case 0x09: // read_int32_info, UnpackUtils.cs:367-382
            if (byte_length != 4) return fail_id(id);
            memcpy(c.int32_info, p, 4);
            c.have_int32 = true;
            break;
        case 0x0d: { // read_channel_info, UnpackUtils.cs:389-410 (the mask over-read is not reproduced; no getter exposes it)
            if (byte_length == 0 || byte_length > 5) return fail_id(id);
            I.num_channels = p[0];
            int64_t mask = 0;
            for (int64_t k = 1; k < byte_length; k++) mask |= (int64_t)p[k] << (8 * (k - 1));
            I.channel_mask = mask;
            break;
        }
        case 0x25: { // read_config_info, UnpackUtils.cs:432-455
            int64_t bytecnt = byte_length;
            int k = 0;
            if (bytecnt >= 3) {
                I.config_flags &= 0xff;
                I.config_flags |= (int64_t)(p[0] << 8);
                I.config_flags |= (int64_t)(p[1] << 16);
                I.config_flags |= (int64_t)(int32_t)((uint32_t)p[2] << 24);
                k = 3;
            }
            if (bytecnt >= 4 && (I.config_flags & 0x2000000) > 0) { I.xmode = p[k++]; bytecnt--; }
            if (bytecnt >= 5) I.five = 1;
            break;
        }
        case 0x27: // read_sample_rate, UnpackUtils.cs:459-473
            if (byte_length == 3) I.sample_rate = (int64_t)p[0] | ((int64_t)p[1] << 8) | ((int64_t)p[2] << 16);
            break;
        case 0x0a: // init_wv_bitstream, UnpackUtils.cs:74-90
            if (!copyable) return fail_id(id);
            B.sub_off[WVB_SUB_WV] = rel; B.sub_len[WVB_SUB_WV] = (uint32_t)byte_length;
            c.wvbits_nonempty = byte_length != 0;
            break;
        case 0x0b: // init_wvc_bitstream, UnpackUtils.cs:96-106 (parsed, never used)
            if ((byte_length & 1) || !copyable) return fail_id(id);
            break;
        case 0x0c:
        case 0x2c: // init_wvx_bitstream, UnpackUtils.cs:115-147
            if (byte_length <= 4 || (byte_length & 1) || !copyable) return fail_id(id);
            B.sub_off[WVB_SUB_WVX] = rel; B.sub_len[WVB_SUB_WVX] = (uint32_t)byte_length;
            if (id == 0x2c) B.bflags |= WVB_BF_WVX_NEW;
            c.wvx_present = true;
            break;
        case 0x0e: { // init_dsd_block, DsdUtils.cs:17-54
            if (byte_length < 2 || p[0] > 31) return fail_id(id);
            if (!copyable) return fail_id(id);
            c.dsd_ready = false;
            I.dsd_multiplier = 1u << p[0];
            int mode = p[1];
            const int64_t n = array_len; // data.Length
            if (mode == 0) {
                if (n - 2 != (int64_t)h.block_samples * ((h.flags & (F_MONO | F_FALSE_STEREO)) ? 1 : 2)) return fail_id(id);
            } else if (mode == 1) {
                if (!dsd_fast_valid(p, (size_t)n, 2)) return fail_id(id);
            } else if (mode == 3) {
                if (n - 2 < ((h.flags & (F_MONO | F_FALSE_STEREO)) ? 13 : 20)) return fail_id(id);
                if (p[3] != 20) return fail_id(id);
            } else
                return fail_id(id);
            c.dsd_ready = true;
            B.sub_off[WVB_SUB_DSD] = rel; B.sub_len[WVB_SUB_DSD] = (uint32_t)n;
            // planner key (see wvb_dsd_core.cuh): mode | history_bits << 4 | rate_i << 8
            B.smem_words = (uint16_t)((mode & 15) | (mode == 1 ? (p[2] & 15) << 4 : 0) | (mode == 3 ? p[2] << 8 : 0));
            if (n != byte_length) B.bflags |= WVB_BF_DSD_PADDED;
            break;
        }
        case 0x2a: // read_new_config_info, UnpackUtils.cs:415-427
            I.five = 1;
            if (byte_length >= 1) I.file_format = p[0];
            break;
        case 0x21:
        case 0x23: // read_header, UnpackUtils.cs:475-482
            if (byte_length < 0) { c.exception = true; return false; }
            I.header_off = (int64_t)data_at; I.header_len = byte_length;
            break;
        case 0x22:
        case 0x24: // read_trailer, UnpackUtils.cs:484-491
            if (byte_length < 0) { c.exception = true; return false; }
            I.trailer_off = (int64_t)data_at; I.trailer_len = byte_length;
            break;
        case 0x28: { // ID_ALT_EXTENSION, MetadataUtils.cs:179-181
            if (byte_length < 0) { c.exception = true; return false; }
            size_t n = (size_t)std::min<int64_t>(byte_length, (int64_t)sizeof(I.file_extension) - 1);
            memcpy(I.file_extension, p, n);
            I.file_extension[n] = 0;
            break;
        }
        case 0x2f: // ID_BLOCK_CHECKSUM: the reference notes it (MetadataUtils.cs:183); the batch decoder verifies it on the device
            I.five = 1;
            if ((byte_length == 2 || byte_length == 4) && bytes_to_read == byte_length && !(rel & 1)) {
                B.bflags |= WVB_BF_BLOCK_CHECKSUM;
                B.checksum_off = rel - 2;
            }
            break;
        default:
            if (!(id & 0x20)) return fail_id(id);
            break;
        }
    }

    if (bytecount != (int64_t)h.ckSize) {
        snprintf(I.error_message, sizeof(I.error_message), "invalid reading WavPack metadata block");
        return false;
    }
    if ((h.block_samples != 0 && (h.flags & F_DSD)) ? !c.dsd_ready : !c.wvbits_nonempty) {
        snprintf(I.error_message, sizeof(I.error_message), "invalid WavPack file");
        return false;
    }
    if (h.block_samples != 0) {
        if ((h.flags & F_INT32) && c.int32_info[0] != 0 && !c.wvx_present) I.lossy_blocks = 1;
        if ((h.flags & F_FLOAT) && (c.float_info[0] & (0x20 | 0x08 | 0x04 | 0x02))) I.lossy_blocks = 1;
    }
    memcpy(B.int32_info, c.int32_info, 4);
    memcpy(B.float_info, c.float_info, 4);
    { // shared-memory words per thread for the decorrelation state (layout: wvb_pcm.cuh) and the grouping signature
        uint32_t words = (uint32_t)c.num_terms, sig = 2166136261u;
        for (int k = 0; k < c.num_terms; k++) {
            int t = c.terms[k];
            if (!stereo && t < 0) t &= 7; // decorr_mono_pass's default branch (UnpackUtils.cs:1207)
            int ring = t > 8 ? 2 : t < 0 ? 1 : t <= 1 ? 1 : t <= 2 ? 2 : t <= 4 ? 4 : 8;
            words += stereo ? 2 + 2 * ring : 1 + ring;
            sig = (sig ^ (uint32_t)(uint8_t)c.terms[k]) * 16777619u;
        }
        if (!(h.flags & F_DSD)) {
            B.smem_words = (uint16_t)words;
            B.terms_sig = sig;
        }
    }
    if (c.have_int32) B.bflags |= WVB_BF_HAS_INT32_INFO;
    if (c.have_float) B.bflags |= WVB_BF_HAS_FLOAT_INFO;
    if (c.wvx_present) B.bflags |= WVB_BF_WVX_PRESENT;
    // state the device cannot rebuild from this block alone (quirk C-8)
    if (h.block_samples != 0 && (h.flags & F_DSD) && !B.sub_off[WVB_SUB_DSD]) {
        // wps.dsd is still the previous block's (exhausted) decoder: mode 0 writes nothing, modes 1/3 emit state-dependent bytes;
        // the device writes zeros, mutes the last piece like the CRC failure would, and flags the block inexact
        B.bflags |= WVB_BF_MUTE_ALL | WVB_BF_STALE_STATE;
        B.smem_words = 0;
    }
    const bool pcm = h.block_samples != 0 && !(h.flags & F_DSD);
    if (pcm) {
        if (!c.terms_from_this_block && c.num_terms > 0) B.bflags |= WVB_BF_STALE_STATE;
        if (!weights_ok) B.bflags |= WVB_BF_STALE_STATE;
        if (!B.sub_off[WVB_SUB_WV]) B.bflags |= WVB_BF_STALE_STATE;      // bitstream carried from an earlier block
        if (!B.sub_off[WVB_SUB_ENTROPY]) B.bflags |= WVB_BF_STALE_STATE; // words_data carried over
        if (c.wvx_present && !B.sub_off[WVB_SUB_WVX] && !(h.flags & F_FLOAT)) B.bflags |= WVB_BF_STALE_STATE; // stale wvxbits/crc_mvx
        if (c.cur_inexact) B.bflags |= WVB_BF_STALE_STATE;
    }
    return true;
}

void finish_open(Ctx &c, uint32_t open_flags) // WavPackUtils.cs:68-117
{
    wvb_file_info &I = *c.info;
    const Header &h = c.hdr;
    I.config_flags = (I.config_flags & ~0xffLL) | (h.flags & 0xff);
    I.bytes_per_sample = (int)((h.flags & F_BYTES_STORED) + 1);
    I.bits_per_sample = I.bytes_per_sample * 8 - (int)((h.flags & SHIFT_MASK) >> SHIFT_LSB);
    if (I.config_flags & F_FLOAT) { I.bytes_per_sample = 3; I.bits_per_sample = 24; }
    if (I.sample_rate == 0) {
        if (h.block_samples == 0 || (h.flags & SRATE_MASK) == SRATE_MASK) I.sample_rate = 44100;
        else I.sample_rate = kSampleRates[(h.flags & SRATE_MASK) >> SRATE_LSB];
    }
    if (I.num_channels == 0) {
        I.num_channels = (h.flags & F_MONO) ? 1 : 2;
        I.channel_mask = 0x5 - I.num_channels;
    }
    if ((open_flags & WVB_OPEN_2CH_MAX) && !(h.flags & F_FINAL)) I.reduced_channels = (h.flags & F_MONO) ? 1 : 2;
    if (!(open_flags & (WVB_OPEN_2CH_MAX | WVB_OPEN_ALL_CHANNELS)) && I.num_channels > 2) {
        snprintf(I.error_message, sizeof(I.error_message), "only two channels supported!");
        I.status = WVB_E_FORMAT;
    }
    if (h.flags & F_DSD) { I.bytes_per_sample = 1; I.bits_per_sample = 8; }
    I.version = h.version;
    I.first_flags = h.flags;
}

struct Sink {
    wvb_block_desc *blocks;
    size_t cap, n = 0;
    bool overflow = false;
    void push(const wvb_block_desc &b)
    {
        if (blocks && n < cap) blocks[n] = b;
        else if (blocks) overflow = true;
        n++;
    }
};

// Reference-faithful sequencing: emulates repeated WavpackUnpackSamples calls until one returns 0.
// Call sizes: `chunk` samples, after an optional prologue of seek()'s skip calls (WavPackUtils.cs:573-578): skip_total
// samples consumed in calls of at most skip_chunk.  max_blocks > 0 ends the walk after that many descriptors.
void index_reference_order(Ctx &c, uint32_t chunk, Sink &out, int64_t skip_total = 0, uint32_t skip_chunk = 0, size_t max_blocks = 0)
{
    wvb_file_info &I = *c.info;
    const int out_ch = I.reduced_channels > 0 ? I.reduced_channels : I.num_channels;
    int64_t out_pos = 0;          // complete samples emitted so far
    int64_t call_remaining = 0;   // samples the current call may still return
    int64_t call_unpacked = 0;
    int64_t skip_left = skip_total; // samples seek()'s skip loop still has to consume (`index`), kept current sample by sample
    bool skip_call = false;       // the current call is one of seek()'s
    bool inited = true;           // the block in c.hdr went through unpack_init (true right after open)
    uint32_t pending_gap = 0;
    bool first_iter = true;
    auto begin_call = [&]() {
        skip_call = skip_left > 0;
        call_remaining = skip_call ? std::min<int64_t>(skip_left, skip_chunk ? skip_chunk : chunk) : (int64_t)chunk;
        call_unpacked = 0;
    };
    for (;;) {
        if (max_blocks && out.n >= max_blocks) break;
        if (call_remaining == 0) {
            if (!first_iter) {
                if (call_unpacked == 0) { // the caller stops on a call that returned 0 (WvDemo.cs:133); seek() would spin forever
                    if (skip_call) I.stopped_early = 1;
                    break;
                }
            }
            begin_call();
        }
        first_iter = false;
        Header &h = c.hdr;
        bool brk = false;
        if (h.block_samples == 0 || !(h.flags & F_INITIAL) || c.sample_index >= h.block_index + (int64_t)h.block_samples) {
            if (!read_next_header(c)) brk = true;
            else if (h.block_samples == 0 || c.sample_index == h.block_index) {
                inited = true;
                if (!unpack_init(c)) { // the call ends here; the next call decodes this header with the state left behind
                    brk = true;
                    I.stopped_early = 1;
                    inited = false;
                }
            } else
                inited = false;
        }
        if (brk) {
            if (c.exception) { I.stopped_early = 1; break; }
            if (call_unpacked == 0) { if (skip_call) I.stopped_early = 1; break; }
            call_remaining = 0; // this call returns short; the caller calls again
            continue;
        }
        if (h.block_samples == 0 || !(h.flags & F_INITIAL) || c.sample_index >= h.block_index + (int64_t)h.block_samples) continue;

        if (c.sample_index < h.block_index) { // zero fill, WavPackUtils.cs:227-251
            int64_t n = std::min<int64_t>(h.block_index - c.sample_index, call_remaining);
            c.sample_index += n;
            call_unpacked += n;
            call_remaining -= n;
            if (skip_call) skip_left -= n;
            pending_gap += (uint32_t)n;
            out_pos += n;
            continue;
        }
        // consume the block call by call; one descriptor for the whole block
        int64_t n_block = h.block_index + (int64_t)h.block_samples - c.sample_index;
        wvb_block_desc B;
        if (inited)
            B = c.cur;
        else { // reached without unpack_init (after a gap): decoder state is whatever the previous block left
            memset(&B, 0, sizeof(B));
            B.in_offset = h.pos; B.in_bytes = (uint32_t)std::min<uint64_t>((uint64_t)h.ckSize + 8, c.len - h.pos); B.block_samples = h.block_samples; B.flags = h.flags; B.crc = h.crc;
            B.block_index = h.block_index; B.version = (uint16_t)h.version;
            B.bflags = WVB_BF_MUTE_ALL;
        }
        if (n_block != (int64_t)h.block_samples) B.bflags |= WVB_BF_MUTE_ALL; // only possible on the non-inited path
        B.block_samples = (uint32_t)n_block;
        B.out_offset = (uint64_t)out_pos;
        B.gap_before = pending_gap;
        pending_gap = 0;
        if (skip_call) { // a block seek() decodes and discards (part of): skip calls consume its head, the caller's grid begins after them
            const uint32_t sc = skip_chunk ? skip_chunk : chunk;
            B.skip_samples = (uint32_t)std::min<int64_t>(skip_left, n_block);
            B.skip_chunk = sc;
            B.chunk_first = chunk;
            // the descriptor's skip grid starts at the block's first sample; a block entered in the middle of a skip call
            // (only after a probe sequence that ran out several blocks early) is cut differently by the reference: flagged
            if (call_remaining != std::min<int64_t>(skip_left, sc)) B.bflags |= WVB_BF_STALE_STATE;
        } else
            B.chunk_first = (uint32_t)std::min<int64_t>(call_remaining, 0xffffffffLL);
        B.avg_block_size = (uint32_t)std::min<int64_t>(h.avg, 0xffffffffLL);
        B.chunk_samples = chunk;
        const bool mono_block = (h.flags & F_MONO) != 0;
        B.out_channels = (uint8_t)(mono_block ? 1 : 2);
        B.out_stride = (uint8_t)out_ch;
        B.out_ch_offset = 0;
        B.out_bps = (uint8_t)I.bytes_per_sample;
        if (B.out_channels > B.out_stride) {
            // a stereo block in a file whose config says one channel: the reference writes two entries per sample into a
            // buffer it advances by one, each piece overwriting half of the previous one.  Not reproduced: zeros, flagged.
            B.out_channels = B.out_stride;
            B.bflags |= WVB_BF_MUTE_ALL | WVB_BF_STALE_STATE;
        }
        out.push(B);
        int64_t left = n_block;
        while (left > 0) {
            if (call_remaining == 0) begin_call();
            int64_t n = std::min(left, call_remaining);
            left -= n; call_remaining -= n; call_unpacked += n;
            if (skip_call) skip_left -= n;
        }
        c.sample_index += n_block;
        out_pos += n_block;
        if (c.sample_index == I.total_samples) call_remaining = 0; // `break` at WavPackUtils.cs:277: the call returns short
    }
    I.indexed_samples = out_pos;
}

// Extension: every audio block of every segment, hop by ckSize+8, channels laid side by side.
void index_all_channels(Ctx &c, uint32_t chunk, Sink &out)
{
    wvb_file_info &I = *c.info;
    int ch_off = 0;
    bool first = true;
    for (;;) {
        if (!first) {
            if (!read_next_header(c)) break;
            if (!unpack_init(c)) { I.stopped_early = 1; break; }
        }
        first = false;
        const Header &h = c.hdr;
        if (h.block_samples == 0) continue;
        if (h.flags & F_INITIAL) ch_off = 0;
        wvb_block_desc B = c.cur;
        B.out_offset = (uint64_t)h.block_index;
        B.chunk_first = 0xffffffffu;
        B.chunk_samples = chunk;
        B.out_channels = (uint8_t)((h.flags & F_MONO) ? 1 : 2);
        B.out_stride = (uint8_t)I.num_channels;
        B.out_ch_offset = (uint8_t)ch_off;
        B.out_bps = (uint8_t)I.bytes_per_sample;
        if (ch_off + B.out_channels > I.num_channels) continue; // more coded channels than the file's config declares: no slot to put them in
        ch_off += B.out_channels;
        out.push(B);
        I.indexed_samples = std::max<int64_t>(I.indexed_samples, h.block_index + (int64_t)h.block_samples);
    }
}

} // namespace

// ---- layout pins: the hand-written mirrors (ctypes in _native.py, [StructLayout] in csharp/WavPackUtils.cs) depend on these
static_assert(sizeof(wvb_block_desc) == 160 && alignof(wvb_block_desc) == 8, "wvb_block_desc layout is part of the ABI");
static_assert(sizeof(wvb_block_result) == 16, "wvb_block_result layout is part of the ABI");
static_assert(sizeof(wvb_seek_state) == 40, "wvb_seek_state layout is part of the ABI");
static_assert(sizeof(wvb_file_info) == 224, "wvb_file_info layout is part of the ABI");
static_assert(offsetof(wvb_block_desc, sub_off) == 40 && offsetof(wvb_block_desc, int32_info) == 104 && offsetof(wvb_block_desc, chunk_first) == 124 &&
              offsetof(wvb_block_desc, skip_samples) == 144, "wvb_block_desc layout is part of the ABI");

extern "C" {

const char *wvb_abi_layout(void)
{
    static std::string text;
    static std::once_flag once;
    std::call_once(once, [] {
        auto add = [&](const char *name, size_t off, size_t size) { text += std::string(name) + ":" + std::to_string(off) + ":" + std::to_string(size) + ";"; };
#define S(T) text += std::string(text.empty() ? "" : "|") + #T + ":" + std::to_string(sizeof(T)) + ";"
#define F(T, f) add(#f, offsetof(T, f), sizeof(((T *)0)->f))
        S(wvb_block_desc);
        F(wvb_block_desc, in_offset); F(wvb_block_desc, out_offset); F(wvb_block_desc, in_bytes); F(wvb_block_desc, block_samples);
        F(wvb_block_desc, flags); F(wvb_block_desc, crc); F(wvb_block_desc, block_index); F(wvb_block_desc, sub_off); F(wvb_block_desc, sub_len);
        F(wvb_block_desc, int32_info); F(wvb_block_desc, float_info); F(wvb_block_desc, bflags); F(wvb_block_desc, version);
        F(wvb_block_desc, out_channels); F(wvb_block_desc, out_stride); F(wvb_block_desc, out_ch_offset); F(wvb_block_desc, out_bps);
        F(wvb_block_desc, smem_words); F(wvb_block_desc, chunk_first); F(wvb_block_desc, chunk_samples); F(wvb_block_desc, file_id);
        F(wvb_block_desc, gap_before); F(wvb_block_desc, terms_sig); F(wvb_block_desc, skip_samples); F(wvb_block_desc, skip_chunk);
        F(wvb_block_desc, avg_block_size); F(wvb_block_desc, checksum_off);
        S(wvb_block_result);
        F(wvb_block_result, crc); F(wvb_block_result, rflags); F(wvb_block_result, mute_from); F(wvb_block_result, crc_x);
        S(wvb_file_info);
        F(wvb_file_info, status); F(wvb_file_info, error_message); F(wvb_file_info, total_samples); F(wvb_file_info, sample_rate);
        F(wvb_file_info, config_flags); F(wvb_file_info, channel_mask); F(wvb_file_info, num_channels); F(wvb_file_info, reduced_channels);
        F(wvb_file_info, bits_per_sample); F(wvb_file_info, bytes_per_sample); F(wvb_file_info, float_norm_exp); F(wvb_file_info, xmode);
        F(wvb_file_info, version); F(wvb_file_info, five); F(wvb_file_info, file_format); F(wvb_file_info, lossy_blocks);
        F(wvb_file_info, dsd_multiplier); F(wvb_file_info, first_flags); F(wvb_file_info, header_off); F(wvb_file_info, header_len);
        F(wvb_file_info, trailer_off); F(wvb_file_info, trailer_len); F(wvb_file_info, file_extension); F(wvb_file_info, num_blocks);
        F(wvb_file_info, indexed_samples); F(wvb_file_info, stopped_early); F(wvb_file_info, reserved);
        S(wvb_seek_state);
        F(wvb_seek_state, hdr_pos); F(wvb_seek_state, block_index); F(wvb_seek_state, avg_block_size); F(wvb_seek_state, file_pos);
        F(wvb_seek_state, block_samples); F(wvb_seek_state, ck_size);
#undef S
#undef F
    });
    return text.c_str();
}

int wvb_index(const uint8_t *file, size_t len, uint32_t open_flags, uint32_t chunk_samples, wvb_file_info *info,
              wvb_block_desc *blocks, size_t cap, size_t *nblocks)
{
    if (!file || !info) return WVB_E_ARG;
    if (chunk_samples == 0) chunk_samples = 4096;
    memset(info, 0, sizeof(*info));
    info->total_samples = -1;
    info->header_len = info->trailer_len = -1;
    if (nblocks) *nblocks = 0;
    Ctx c;
    c.d = file; c.len = len; c.info = info;
    memset(&c.cur, 0, sizeof(c.cur));
    // WavpackOpenFileInput, WavPackUtils.cs:47-66
    while (c.hdr.block_samples == 0) {
        if (!read_next_header(c)) {
            snprintf(info->error_message, sizeof(info->error_message), "not compatible with this version of WavPack file!");
            info->status = WVB_E_FORMAT;
            return WVB_OK;
        }
        if (c.hdr.block_samples > 0 && c.hdr.total_samples != 0xFFFFFFFFLL) info->total_samples = c.hdr.total_samples;
        if (!unpack_init(c)) {
            if (c.exception) snprintf(info->error_message, sizeof(info->error_message), "exception");
            info->status = WVB_E_FORMAT;
            return WVB_OK;
        }
    }
    finish_open(c, open_flags);
    if (info->status != WVB_OK) return WVB_OK;
    Sink sink{blocks, cap};
    if (open_flags & WVB_OPEN_ALL_CHANNELS) index_all_channels(c, chunk_samples, sink);
    else index_reference_order(c, chunk_samples, sink);
    info->num_blocks = (int64_t)sink.n;
    if (nblocks) *nblocks = sink.n;
    if (sink.overflow) return WVB_E_CAPACITY;
    return WVB_OK;
}

// seek() (WavPackUtils.cs:521-594): where does the reference restart its decoder?  Replays its probe sequence on the headers
// (from != NULL) or hops from header to header (from == NULL).  Returns false where seek() returns false.
static bool seek_find_block(const uint8_t *file, size_t len, const wvb_seek_state *from, int64_t target, size_t *block_pos, int64_t *index)
{
    Ctx r;
    wvb_file_info scratch;
    r.d = file; r.len = len; r.info = &scratch;
    if (!from) {
        while (read_next_header(r)) {
            const Header &h = r.hdr;
            if (h.block_samples != 0 && target >= h.block_index && target < h.block_index + (int64_t)h.block_samples) {
                *block_pos = h.pos;
                *index = target - h.block_index;
                return true;
            }
            r.pos = std::min<size_t>(len, std::max<size_t>(h.pos + (size_t)h.ckSize + 8, h.pos + 32));
        }
        return false;
    }
    Header &h = r.hdr;
    h.pos = (size_t)from->hdr_pos; h.block_index = from->block_index; h.block_samples = from->block_samples; h.ckSize = from->ck_size;
    h.avg = from->avg_block_size;
    r.pos = (size_t)std::min<int64_t>(std::max<int64_t>(from->file_pos, 0), (int64_t)len);
    int steps = 25;     // "maximum steps to position"
    const int near_blocks = 5; // closer than this: walk the headers forward instead of jumping
    while (steps-- > 0) {
        int64_t seek_pos = (int64_t)h.pos;
        if (target <= (int64_t)h.block_samples)
            seek_pos = 0;
        else if (target < h.block_index || target > h.block_index + (int64_t)h.block_samples) {
            int64_t distance = target - h.block_index;
            distance += distance > 0 ? 1 - (int64_t)h.block_samples : 1 - 2 * (int64_t)h.block_samples; // back-off so that the walk ends going forward
            if (h.block_samples == 0) return false; // the reference divides by zero here (uncaught): no seek either way
            const int64_t blocks = distance / (int64_t)h.block_samples;
            if (blocks >= 0 && blocks <= near_blocks) seek_pos = -1;
            else seek_pos += blocks * h.avg;
            if (seek_pos >= (int64_t)len) seek_pos = -1;
        }
        if (seek_pos != -1) {
            if (seek_pos < 0) return false; // Stream.Seek before the start: IOException, caught, `return false`
            r.pos = (size_t)seek_pos;
        }
        if (!read_next_header(r)) continue; // wphdr.error: the step is spent, the header fields stay
        if (steps == 0 || (target >= h.block_index && target < h.block_index + (int64_t)h.block_samples)) {
            *block_pos = h.pos;
            *index = target - h.block_index;
            return true;
        }
        if (seek_pos == -1) {
            r.pos = std::min<size_t>(len, h.pos + (size_t)h.ckSize);
            steps--; // the reference means to not count header walks and decrements once more instead
        }
    }
    return false;
}

int wvb_index_seek(const uint8_t *file, size_t len, uint32_t open_flags, const wvb_seek_state *from, int64_t target, uint32_t skip_chunk,
                   uint32_t chunk_samples, size_t max_blocks, wvb_file_info *info, wvb_block_desc *blocks, size_t cap, size_t *nblocks,
                   int64_t *window_first_sample, int64_t *landed_sample)
{
    if (!file || !info) return WVB_E_ARG;
    if (open_flags & WVB_OPEN_ALL_CHANNELS) return WVB_E_ARG; // the extension has its own order; seek it by block_index in the full table
    if (chunk_samples == 0) chunk_samples = 4096;
    if (nblocks) *nblocks = 0;
    if (window_first_sample) *window_first_sample = 0;
    if (landed_sample) *landed_sample = 0;
    size_t n_all = 0;
    int rc = wvb_index(file, len, open_flags, chunk_samples, info, nullptr, 0, &n_all); // the file as WavpackOpenFileInput sees it
    if (rc != WVB_OK || info->status != WVB_OK) return rc;
    info->num_blocks = 0;
    info->indexed_samples = 0;
    info->stopped_early = 0;
    if (target >= info->total_samples) return WVB_OK;   // WavPackUtils.cs:527-528 (also: unknown length, total_samples == -1)
    if (target < 0) target = 0;                         // WavPackUtils.cs:529-530
    size_t block_pos = 0;
    int64_t index = 0;
    if (!seek_find_block(file, len, from, target, &block_pos, &index)) return WVB_OK;
    // `WavpackContext c = WavpackOpenFileInput(infile); wpc.stream = c.stream;` (WavPackUtils.cs:568-570): a fresh stream state
    // at that block; the nested context's config is dropped, the caller's (info) stays
    wvb_file_info scratch = *info;
    Ctx c;
    c.d = file; c.len = len; c.info = &scratch;
    c.pos = block_pos;
    memset(&c.cur, 0, sizeof(c.cur));
    while (c.hdr.block_samples == 0) {
        if (!read_next_header(c)) return WVB_OK;
        if (!unpack_init(c)) { info->stopped_early = 1; return WVB_OK; }
    }
    const int out_ch = info->reduced_channels > 0 ? info->reduced_channels : (info->num_channels ? info->num_channels : 2);
    if (skip_chunk == 0) skip_chunk = (uint32_t)(4096 / out_ch); // Defines.SAMPLE_BUFFER_SIZE / WavpackGetReducedChannels
    const int64_t first = c.sample_index; // the restarted decoder's position: the block's block_index
    Sink sink{blocks, cap};
    index_reference_order(c, chunk_samples, sink, std::max<int64_t>(index, 0), skip_chunk, max_blocks);
    info->num_blocks = (int64_t)sink.n;
    info->indexed_samples = scratch.indexed_samples;
    info->stopped_early = scratch.stopped_early;
    info->lossy_blocks |= scratch.lossy_blocks;
    if (nblocks) *nblocks = sink.n;
    if (window_first_sample) *window_first_sample = first;
    if (landed_sample) *landed_sample = first + std::max<int64_t>(index, 0); // index <= 0: nothing is skipped, the reader stands at the block's start
    if (sink.overflow) return WVB_E_CAPACITY;
    return WVB_OK;
}

// ID_MD5_CHECKSUM (Defines.cs:77) lookup: a plain walk over every block's sub-blocks (TLV layout: MetadataUtils.cs:25-82),
// independent of the reference's read order (the reference never looks at this id).
int wvb_stored_md5(const uint8_t *file, size_t len, uint8_t md5[16])
{
    if (!file || !md5) return 0;
    size_t pos = 0;
    while (pos + 32 <= len) {
        if (memcmp(file + pos, "wvpk", 4) != 0) { pos++; continue; }
        const uint32_t ck = (uint32_t)file[pos + 4] | ((uint32_t)file[pos + 5] << 8) | ((uint32_t)file[pos + 6] << 16) | ((uint32_t)file[pos + 7] << 24);
        if (ck < 24 || ck >= 0x100000u || (ck & 1) || pos + 8 + (size_t)ck > len) { pos++; continue; }
        size_t at = pos + 32;
        const size_t end = pos + 8 + ck;
        while (at + 2 <= end) {
            const uint8_t id = file[at];
            size_t words = file[at + 1];
            size_t hdr = 2;
            if (id & 0x80) {
                if (at + 4 > end) break;
                words |= ((size_t)file[at + 2] << 8) | ((size_t)file[at + 3] << 16);
                hdr = 4;
            }
            const size_t padded = words * 2;
            if (at + hdr + padded > end) break;
            const size_t actual = (id & 0x40) ? (padded ? padded - 1 : 0) : padded;
            if ((id & 0x3f) == 0x26 && actual == 16) {
                memcpy(md5, file + at + hdr, 16);
                return 1;
            }
            at += hdr + padded;
        }
        pos = end;
    }
    return 0;
}

int wvb_block_checksum_ok(const uint8_t *block, size_t len)
{
    if (!block || len < 32 || memcmp(block, "wvpk", 4) != 0) return -1;
    const size_t end = std::min<size_t>(len, (size_t)le32(block + 4) + 8);
    size_t at = 32;
    while (at + 2 <= end) {
        const uint8_t id = block[at];
        size_t words = block[at + 1], hdr = 2;
        if (id & 0x80) {
            if (at + 4 > end) return -1;
            words |= ((size_t)block[at + 2] << 8) | ((size_t)block[at + 3] << 16);
            hdr = 4;
        }
        if (at + hdr + 2 * words > end) return -1;
        if ((id & 0x3f) == 0x2f) {
            const size_t n = 2 * words;
            if ((id & 0x40) || hdr != 2 || (n != 2 && n != 4) || (at & 1)) return -1;
            uint32_t csum = 0xffffffffu;
            for (size_t i = 0; i < at; i += 2) csum = csum * 3u + ((uint32_t)block[i] | ((uint32_t)block[i + 1] << 8));
            const uint8_t *st = block + at + 2;
            if (n == 4) return le32(st) == csum;
            csum ^= csum >> 16;
            return ((uint32_t)st[0] | ((uint32_t)st[1] << 8)) == (csum & 0xffffu);
        }
        at += hdr + 2 * words;
    }
    return -1;
}

uint32_t wvb_frame_bytes(const wvb_block_desc *b, int out_format)
{
    uint32_t unit = out_format == WVB_OUT_INT32 ? 4u : b->out_bps;
    return unit * b->out_stride;
}

void wvb_rebase(wvb_block_desc *blocks, size_t n, uint64_t in_base, uint64_t out_base, int out_format, uint32_t file_id)
{
    for (size_t i = 0; i < n; i++) {
        wvb_block_desc &b = blocks[i];
        b.in_offset += in_base;
        b.out_offset = out_base + b.out_offset * wvb_frame_bytes(&b, out_format);
        b.file_id = file_id;
    }
}

int wvb_index_many(const uint8_t *slab, const uint64_t *offsets, const uint64_t *sizes, size_t nfiles, uint32_t open_flags,
                   uint32_t chunk_samples, int out_format, int threads, wvb_file_info *infos, wvb_block_desc *blocks, size_t cap,
                   uint64_t *first, uint64_t *count, uint64_t *file_out_offset, size_t *nblocks, uint64_t *out_bytes)
{
    if (!slab || !offsets || !sizes || !infos || !first || !count) return WVB_E_ARG;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(nfiles, 1));
    // One walk over the files.  With a table to fill, every worker collects the descriptors of the files it takes in a
    // buffer of its own (files are handed out dynamically, so the final position of a file's descriptors is only known
    // once every file before it has been counted); a second, copy-only step moves them into the dense, file-ordered table.
    struct Piece { int worker; size_t start; };
    std::vector<Piece> piece(blocks ? nfiles : 0);
    std::vector<std::vector<wvb_block_desc>> local(blocks ? (size_t)threads : 0);
    std::atomic<size_t> next{0};
    auto walk = [&](int w) {
        std::vector<wvb_block_desc> *mine = blocks ? &local[(size_t)w] : nullptr;
        if (mine) mine->reserve(cap / (size_t)threads + 64);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nfiles) break;
            size_t n = 0;
            if (!mine) {
                wvb_index(slab + offsets[i], (size_t)sizes[i], open_flags, chunk_samples, &infos[i], nullptr, 0, &n);
            } else {
                // grow-and-retry: the file's block count is not known before the walk (rarely more than one retry per worker)
                const size_t start = mine->size();
                size_t room = std::max<size_t>(mine->capacity() - start, 64);
                for (;;) {
                    mine->resize(start + room);
                    const int rc = wvb_index(slab + offsets[i], (size_t)sizes[i], open_flags, chunk_samples, &infos[i], mine->data() + start, room, &n);
                    if (rc != WVB_E_CAPACITY) break;
                    room = n;
                }
                mine->resize(start + n);
                piece[i] = Piece{w, start};
            }
            count[i] = n;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(walk, t);
        for (auto &t : th) t.join();
    }
    uint64_t total = 0, obytes = 0;
    std::vector<uint64_t> own_offsets; // the caller may not want the per-file offsets; the rebase needs them either way
    if (!file_out_offset) {
        own_offsets.resize(nfiles);
        file_out_offset = own_offsets.data();
    }
    for (size_t i = 0; i < nfiles; i++) {
        first[i] = total;
        total += count[i];
        uint32_t unit = out_format == WVB_OUT_INT32 ? 4u : (uint32_t)infos[i].bytes_per_sample;
        uint32_t ch = (open_flags & WVB_OPEN_ALL_CHANNELS) ? (uint32_t)infos[i].num_channels
                                                           : (uint32_t)(infos[i].reduced_channels > 0 ? infos[i].reduced_channels : infos[i].num_channels);
        file_out_offset[i] = obytes;
        obytes += (uint64_t)infos[i].indexed_samples * unit * ch;
        obytes = (obytes + 15) & ~(uint64_t)15;
    }
    if (nblocks) *nblocks = (size_t)total;
    if (out_bytes) *out_bytes = obytes;
    if (!blocks) return WVB_OK;
    if (total > cap) return WVB_E_CAPACITY;
    next = 0;
    auto gather = [&]() {
        for (;;) {
            const size_t i0 = next.fetch_add(64);
            if (i0 >= nfiles) break;
            for (size_t i = i0; i < std::min(nfiles, i0 + 64); i++) {
                const size_t n = (size_t)count[i];
                if (!n) continue;
                memcpy(blocks + first[i], local[(size_t)piece[i].worker].data() + piece[i].start, n * sizeof(wvb_block_desc));
                wvb_rebase(blocks + first[i], n, offsets[i], file_out_offset[i], out_format, (uint32_t)i);
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(gather);
        for (auto &t : th) t.join();
    }
    return WVB_OK;
}

} // extern "C"
