/*
 * oracle/refdec.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see refdec.h).
 *
 * CPU restatement of the reference's block decode path.  Integer semantics are
 * the C# ones: 32-bit int arithmetic wraps (build with -fwrapv), shift counts
 * are masked (int: &31, long: &63), `long` is 64-bit.  Array accesses that would
 * raise IndexOutOfRangeException in C# longjmp to the API entry ("exception").
 *
 * Parity status: UNPINNED by reference-owned vectors (the reference has none and
 * cannot run here); cross-pinned against FFmpeg's independent implementation,
 * see tests/golden/README.md.
 */
#include "refdec.h"

#include <math.h>
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* C# semantics helpers                                                */
/* ------------------------------------------------------------------ */
typedef int32_t i32;
typedef uint32_t u32;
typedef int64_t i64;
typedef uint64_t u64;

static inline i32 shl32(i32 x, int n) { return (i32)((u32)x << (n & 31)); }
static inline i32 sar32(i32 x, int n) { return x >> (n & 31); }
static inline u32 shlu32(u32 x, int n) { return x << (n & 31); }
static inline u32 shru32(u32 x, int n) { return x >> (n & 31); }
static inline i64 shl64(i64 x, int n) { return (i64)((u64)x << (n & 63)); }
static inline i64 sar64(i64 x, int n) { return x >> (n & 63); }

/* Defines.cs:28-94 */
enum {
    BYTES_STORED = 3, MONO_FLAG = 4, HYBRID_FLAG = 8, JOINT_STEREO = 0x10,
    FLOAT_DATA = 0x80, INT32_DATA = 0x100, HYBRID_BITRATE = 0x200, HYBRID_BALANCE = 0x400,
    INITIAL_BLOCK = 0x800, FINAL_BLOCK = 0x1000, SHIFT_LSB = 13, MAG_LSB = 18, SRATE_LSB = 23,
    FALSE_STEREO = 0x40000000, MONO_DATA = MONO_FLAG | FALSE_STEREO,
    MAX_NTERMS = 16, MAX_TERM = 8, BITSTREAM_BUFFER_SIZE = 16 * 1024,
    MIN_STREAM_VERS = 0x402, MAX_STREAM_VERS = 0x410
};
#define DSD_FLAG 0x80000000u
#define SHIFT_MASK (0x1fL << SHIFT_LSB)
#define MAG_MASK (0x1fL << MAG_LSB)
#define SRATE_MASK (0xfL << SRATE_LSB)

enum { /* Defines.cs:50-83 */
    ID_OPTIONAL_DATA = 0x20, ID_ODD_SIZE = 0x40, ID_LARGE = 0x80,
    ID_DUMMY = 0, ID_DECORR_TERMS = 2, ID_DECORR_WEIGHTS = 3, ID_DECORR_SAMPLES = 4, ID_ENTROPY_VARS = 5,
    ID_HYBRID_PROFILE = 6, ID_SHAPING_WEIGHTS = 7, ID_FLOAT_INFO = 8, ID_INT32_INFO = 9, ID_WV_BITSTREAM = 0xa,
    ID_WVC_BITSTREAM = 0xb, ID_WVX_BITSTREAM = 0xc, ID_CHANNEL_INFO = 0xd, ID_DSD_BLOCK = 0xe,
    ID_RIFF_HEADER = 0x21, ID_RIFF_TRAILER = 0x22, ID_ALT_HEADER = 0x23, ID_ALT_TRAILER = 0x24,
    ID_CONFIG_BLOCK = 0x25, ID_MD5_CHECKSUM = 0x26, ID_SAMPLE_RATE = 0x27, ID_ALT_EXTENSION = 0x28,
    ID_NEW_CONFIG_BLOCK = 0x2a, ID_WVX_NEW_BITSTREAM = 0x2c, ID_BLOCK_CHECKSUM = 0x2f
};
enum { /* Defines.cs:96-101 */
    FLOAT_SHIFT_SAME = 2, FLOAT_SHIFT_SENT = 4, FLOAT_ZEROS_SENT = 8, FLOAT_EXCEPTIONS = 0x20
};
/* Defines.cs:112-145 */
#define CONFIG_HYBRID_FLAG 8L
#define CONFIG_FLOAT_DATA 0x80L
#define CONFIG_FAST_FLAG 0x200L
#define CONFIG_HIGH_FLAG 0x800L
#define CONFIG_VERY_HIGH_FLAG 0x1000L
#define CONFIG_LOSSY_MODE 0x1000000L
#define CONFIG_EXTRA_MODE 0x2000000L
enum { MODE_LOSSLESS = 2, MODE_HYBRID = 4, MODE_FLOAT = 8, MODE_HIGH = 0x20, MODE_FAST = 0x40, MODE_EXTRA = 0x80,
       MODE_VERY_HIGH = 0x400, MODE_XMODE = 0x7000, MODE_DSD = 0x10000 };

/* ------------------------------------------------------------------ */
/* State types                                                         */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; int len; int refs; } ByteArray; /* a C# byte[]; len == .Length */

typedef struct { /* Bitstream.cs:15-21 */
    int end, ptr;
    u32 sr;
    int file_bytes;
    int error, bc;
    ByteArray *buf;
    int buf_index;
    int is_null; /* models a null reference */
} Bitstream;

typedef struct { /* decorr_pass.cs:24-26 */
    int16_t term, delta, weight_A, weight_B;
    i32 samples_A[MAX_TERM], samples_B[MAX_TERM];
} decorr_pass;

typedef struct { i32 slow_level; i32 median[3]; i32 error_limit; } entropy_data; /* entropy_data.cs:15-17 */

typedef struct { /* words_data.cs:23-30 */
    i64 bitrate_delta[2], bitrate_acc[2];
    i64 zeros_acc;
    int holding_one, holding_zero;
    entropy_data c[2];
} words_data;

typedef struct { /* WavpackHeader.cs:15-22 */
    u32 ckSize;
    int16_t version;
    i64 total_samples, block_index;
    u32 block_samples, flags;
    i32 crc;
    int error;
    i64 stream_position;
    i64 average_block_size;
} WavpackHeader;

typedef struct { i32 value, filter0, filter1, filter2, filter3, filter4, filter5, filter6, factor; i32 bytei; } DSDfilters;

typedef struct { /* WavpackStream.cs:21-35 */
    ByteArray *data;
    int byteptr;
    uint8_t *probabilities; int probabilities_len;
    uint8_t *lookup_buffer; int lookup_len;
    i32 *value_lookup;
    uint8_t mode;
    int ready;
    int history_bins, p0, p1;
    uint16_t *summed_probabilities;
    u32 low, high, value;
    DSDfilters filters[2]; int filters_present;
    i32 *ptable;
} dsds;

typedef struct { /* WavpackStream.cs:54-62 */
    WavpackHeader wphdr;
    Bitstream wvbits, wvcbits, wvxbits;
    words_data w;
    int num_terms;
    int mute_error;
    i32 crc, crc_x, crc_mvx;
    i64 sample_index;
    int16_t int32_sent_bits, int32_zeros, int32_ones, int32_dups;
    int16_t float_flags, float_shift, float_max_exp, float_norm_exp;
    uint8_t int32_max_width;
    uint8_t float_min_shifted_zeros, float_max_shifted_ones;
    decorr_pass decorr_passes[MAX_NTERMS];
    dsds dsd;
} WavpackStream;

typedef struct { /* WavpackConfig.cs:15-18 */
    int bits_per_sample, bytes_per_sample;
    int num_channels, float_norm_exp;
    i64 flags, sample_rate, channel_mask;
    uint8_t xmode;
} WavpackConfig;

typedef struct { /* WavpackMetadata.cs:15-23 */
    int byte_length;
    ByteArray *data;
    uint8_t id;
    int hasdata, error;
    i64 bytecount;
} WavpackMetadata;

typedef struct { const uint8_t *data; i64 len, pos; } MemStream; /* the caller's BinaryReader/Stream */

struct rd_context { /* WavpackContext.cs:15-35 */
    WavpackConfig config;
    WavpackStream stream;
    ByteArray read_buffer;
    const char *error_message;
    char error_buf[96];
    MemStream infile;
    i64 total_samples, crc_errors;
    int open_flags, norm_offset;
    int reduced_channels;
    int lossy_blocks;
    int five;
    int file_format;
    char *file_extension;
    uint8_t *header; long header_len;
    uint8_t *trailer; long trailer_len;
    u32 dsd_multiplier;
    WavpackMetadata md; /* unpack_init's local wpmd */
    jmp_buf jb; /* "exception" target */
};

#define THROW(c) longjmp((c)->jb, 1)

/* ------------------------------------------------------------------ */
/* Tables (WordsUtils.cs:33-66): closed forms, see tools/gen_tables.py  */
/* ------------------------------------------------------------------ */
static int nbits_table[256], log2_table[256], exp2_table[256], ones_count_table[256];
static int tables_ready;
static void init_tables(void)
{
    if (tables_ready) return;
    for (int i = 0; i < 256; i++) {
        int n = 0, v = i;
        while (v) { n++; v >>= 1; }
        nbits_table[i] = n;
        n = 0; v = i;
        while (v & 1) { n++; v >>= 1; }
        ones_count_table[i] = n;
        log2_table[i] = (int)floor(256.0 * log2(1.0 + i / 256.0) + 0.5);
        exp2_table[i] = (int)floor(256.0 * (pow(2.0, i / 256.0) - 1.0) + 0.5);
    }
    tables_ready = 1;
}
/* exported for tests: which=0 log2, 1 exp2, 2 nbits, 3 ones_count */
int rd_dbg_table(int which, int i)
{
    init_tables();
    i &= 255;
    return which == 0 ? log2_table[i] : which == 1 ? exp2_table[i] : which == 2 ? nbits_table[i] : ones_count_table[i];
}

static const i64 sample_rates[] = { 6000, 8000, 9600, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 64000, 88200, 96000, 192000 }; /* WavPackUtils.cs:18 */

/* ------------------------------------------------------------------ */
/* byte[] model                                                        */
/* ------------------------------------------------------------------ */
static ByteArray *ba_new(int len)
{
    ByteArray *a = (ByteArray *)malloc(sizeof(ByteArray));
    a->p = (uint8_t *)calloc((size_t)(len > 0 ? len : 1), 1);
    a->len = len;
    a->refs = 1;
    return a;
}
static void ba_ref(ByteArray *a) { if (a && a->refs > 0) a->refs++; }
static void ba_unref(ByteArray *a)
{
    if (!a || a->refs <= 0) return; /* refs<=0: statically owned (read_buffer) */
    if (--a->refs == 0) { free(a->p); free(a); }
}
static inline uint8_t ba_get(rd_context *c, ByteArray *a, i64 i)
{
    if (!a || i < 0 || i >= a->len) THROW(c);
    return a->p[i];
}

/* ------------------------------------------------------------------ */
/* Stream model                                                        */
/* ------------------------------------------------------------------ */
static int ms_read(MemStream *s, uint8_t *dst, int cnt) /* Stream.Read: short only at EOF */
{
    i64 left = s->len - s->pos;
    if (left < 0) left = 0;
    if (cnt > left) cnt = (int)left;
    if (cnt > 0) memcpy(dst, s->data + s->pos, (size_t)cnt);
    s->pos += cnt;
    return cnt;
}

/* ------------------------------------------------------------------ */
/* BitsUtils.cs                                                        */
/* ------------------------------------------------------------------ */
static void bs_read(rd_context *c, Bitstream *bs) /* BitsUtils.cs:95-146, file_bytes is always 0 here (bs_open_read passed=0) */
{
    (void)c;
    bs->error = 1;
    memset(bs->buf->p, 0xff, (size_t)bs->buf->len);
    bs->ptr = 0;
    bs->buf_index = 0;
}

static int getbit(rd_context *c, Bitstream *bs) /* BitsUtils.cs:15-35 */
{
    if (bs->bc > 0)
        bs->bc--;
    else {
        bs->ptr++;
        bs->buf_index++;
        bs->bc = 7;
        if (bs->ptr == bs->end)
            bs_read(c, bs);
        bs->sr = ba_get(c, bs->buf, bs->buf_index);
    }
    int result = (bs->sr & 1) > 0;
    bs->sr >>= 1;
    return result;
}

static i64 getbits(rd_context *c, int nbits, Bitstream *bs) /* BitsUtils.cs:37-68 */
{
    i64 retval;
    while (nbits > bs->bc) {
        bs->ptr++;
        bs->buf_index++;
        if (bs->ptr == bs->end)
            bs_read(c, bs);
        bs->sr |= (u32)shl32((i32)ba_get(c, bs->buf, bs->buf_index), bs->bc);
        bs->bc += 8;
    }
    retval = bs->sr;
    if (bs->bc > 32) {
        bs->bc -= nbits;
        bs->sr = (u32)sar32((i32)ba_get(c, bs->buf, bs->buf_index), 8 - bs->bc);
    } else {
        bs->bc -= nbits;
        bs->sr = shru32(bs->sr, nbits);
    }
    return retval;
}

static void bs_open_read(Bitstream *bs, ByteArray *stream, int buffer_start, int buffer_end) /* BitsUtils.cs:70-93, passed==0 */
{
    if (!bs->is_null) ba_unref(bs->buf);
    memset(bs, 0, sizeof(*bs));
    bs->buf = stream;
    ba_ref(stream);
    bs->buf_index = buffer_start;
    bs->end = buffer_end;
    bs->sr = 0;
    bs->bc = 0;
    bs->buf_index--;
    bs->ptr = -1;
}

/* ------------------------------------------------------------------ */
/* WordsUtils.cs: math helpers                                         */
/* ------------------------------------------------------------------ */
static i32 exp2s(i32 log) /* WordsUtils.cs:633-646 */
{
    i64 value;
    if (log < 0)
        return -exp2s(-log);
    value = exp2_table[log & 0xff] | 0x100;
    if ((log >>= 8) <= 9)
        return (i32)sar64(value, 9 - log);
    else
        return (i32)shl64(value, log - 9);
}

static int tbl_nbits(rd_context *c, i64 idx)
{
    if (idx < 0 || idx > 255) THROW(c);
    return nbits_table[idx];
}

static int count_bits(rd_context *c, i64 av) /* WordsUtils.cs:513-537 */
{
    if (av < 256) return tbl_nbits(c, av);
    if (av < 65536) return tbl_nbits(c, av >> 8) + 8;
    if (av < 16777216) return tbl_nbits(c, av >> 16) + 16;
    return tbl_nbits(c, av >> 24) + 24;
}

static int mylog2(rd_context *c, i64 avalue) /* WordsUtils.cs:588-608 */
{
    int dbits;
    if ((avalue += (avalue >> 9)) < (1 << 8)) {
        dbits = tbl_nbits(c, (i32)avalue);
        return (dbits << 8) + log2_table[(i32)shl64(avalue, 9 - dbits) & 0xff];
    } else {
        if (avalue < (1LL << 16))
            dbits = tbl_nbits(c, (i32)(avalue >> 8)) + 8;
        else if (avalue < (1LL << 24))
            dbits = tbl_nbits(c, (i32)(avalue >> 16)) + 16;
        else
            dbits = tbl_nbits(c, (i32)(avalue >> 24)) + 24;
        return (dbits << 8) + log2_table[(i32)sar64(avalue, dbits - 9) & 0xff];
    }
}

static int restore_weight(int8_t weight) /* WordsUtils.cs:653-661 */
{
    int result;
    if ((result = (int)weight << 3) > 0)
        result += (result + 64) >> 7;
    return result;
}

/* ------------------------------------------------------------------ */
/* MetadataUtils.cs / WavpackMetadata.cs                               */
/* ------------------------------------------------------------------ */
static void md_set_data(WavpackMetadata *m, ByteArray *a)
{
    if (m->data == a) return;
    ba_unref(m->data);
    m->data = a;
    ba_ref(a);
}

static int copy_data(WavpackMetadata *m) /* WavpackMetadata.cs:25-36 */
{
    if (!m->hasdata || m->byte_length <= 0) return 0;
    if (m->data->len != BITSTREAM_BUFFER_SIZE) return 1;
    ByteArray *n = ba_new(m->byte_length);
    memcpy(n->p, m->data->p, (size_t)m->byte_length);
    md_set_data(m, n);
    ba_unref(n);
    return 1;
}

static int read_metadata_buff(rd_context *wpc, WavpackMetadata *wpmd) /* MetadataUtils.cs:15-109 */
{
    uint8_t two[2], tchar;
    if (wpmd->bytecount >= (i64)wpc->stream.wphdr.ckSize)
        return 0;
    /* two ReadByte()s: the first may succeed and advance before the second throws */
    if (ms_read(&wpc->infile, &two[0], 1) != 1) { wpmd->error = 1; return 0; }
    wpmd->id = two[0];
    if (ms_read(&wpc->infile, &two[1], 1) != 1) { wpmd->error = 1; return 0; }
    tchar = two[1];
    wpmd->bytecount += 2;
    wpmd->byte_length = tchar << 1;
    if (wpmd->id & ID_LARGE) {
        wpmd->id &= (uint8_t)~ID_LARGE;
        if (ms_read(&wpc->infile, &tchar, 1) != 1) { wpmd->error = 1; return 0; }
        wpmd->byte_length += tchar << 9;
        if (ms_read(&wpc->infile, &tchar, 1) != 1) { wpmd->error = 1; return 0; }
        wpmd->byte_length += tchar << 17;
        wpmd->bytecount += 2;
    }
    int bytes_to_read = wpmd->byte_length;
    if (wpmd->id & ID_ODD_SIZE) {
        wpmd->id &= (uint8_t)~ID_ODD_SIZE;
        wpmd->byte_length--;
    }
    if (wpmd->byte_length == 0) {
        wpmd->hasdata = 0;
        return 1;
    }
    wpmd->bytecount += bytes_to_read;
    if (bytes_to_read > 0) {
        md_set_data(wpmd, &wpc->read_buffer);
        if (bytes_to_read > wpmd->data->len) {
            ByteArray *n = ba_new(bytes_to_read);
            md_set_data(wpmd, n);
            ba_unref(n);
        }
        if (ms_read(&wpc->infile, wpmd->data->p, bytes_to_read) != bytes_to_read) {
            wpmd->hasdata = 0;
            return 0;
        }
        wpmd->hasdata = 1;
    }
    return 1;
}

/* ---- UnpackUtils.cs metadata readers ---- */
static int init_wv_bitstream(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:74-90 */
{
    WavpackStream *wps = &wpc->stream;
    if (!copy_data(wpmd)) return 0;
    bs_open_read(&wps->wvbits, wpmd->data, 0, wpmd->byte_length);
    return 1;
}

static int init_wvc_bitstream(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:96-106 */
{
    WavpackStream *wps = &wpc->stream;
    if ((wpmd->byte_length & 1) > 0 || !copy_data(wpmd)) return 0;
    bs_open_read(&wps->wvcbits, wpmd->data, 0, wpmd->byte_length);
    return 1;
}

static int init_wvx_bitstream(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:115-147 */
{
    WavpackStream *wps = &wpc->stream;
    int counter = 0;
    if (wpmd->byte_length <= 4 || (wpmd->byte_length & 1) > 0 || !copy_data(wpmd)) return 0;
    wps->crc_mvx = ba_get(wpc, wpmd->data, counter++);
    wps->crc_mvx |= ba_get(wpc, wpmd->data, counter++) << 8;
    wps->crc_mvx |= ba_get(wpc, wpmd->data, counter++) << 16;
    wps->crc_mvx |= shl32(ba_get(wpc, wpmd->data, counter++), 24);
    bs_open_read(&wps->wvxbits, wpmd->data, counter, wpmd->byte_length);
    if (wpmd->id == ID_WVX_NEW_BITSTREAM) {
        if ((wps->wphdr.flags & FLOAT_DATA) > 0) {
            wps->float_min_shifted_zeros = (uint8_t)(getbits(wpc, 5, &wps->wvxbits) & 0x1f);
            wps->float_max_shifted_ones = (uint8_t)(getbits(wpc, 5, &wps->wvxbits) & 0x1f);
        } else
            wps->int32_max_width = (uint8_t)(getbits(wpc, 5, &wps->wvxbits) & 0x1f);
    }
    return 1;
}

static int read_decorr_terms(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:156-187 */
{
    WavpackStream *wps = &wpc->stream;
    int termcnt = wpmd->byte_length;
    decorr_pass tmp[MAX_NTERMS]; /* tmpwps = new WavpackStream(): fresh zeroed passes */
    int counter = 0, dcounter;
    if (termcnt > MAX_NTERMS) return 0;
    memset(tmp, 0, sizeof(tmp));
    for (dcounter = termcnt - 1; dcounter >= 0; dcounter--) {
        uint8_t b = ba_get(wpc, wpmd->data, counter);
        tmp[dcounter].term = (int16_t)((int)(b & 0x1f) - 5);
        tmp[dcounter].delta = (int16_t)((b >> 5) & 0x7);
        counter++;
        if (tmp[dcounter].term < -3 || (tmp[dcounter].term > MAX_TERM && tmp[dcounter].term < 17) || tmp[dcounter].term > 18)
            return 0;
    }
    memcpy(wps->decorr_passes, tmp, sizeof(tmp));
    wps->num_terms = termcnt;
    return 1;
}

static int read_decorr_weights(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:196-239 */
{
    WavpackStream *wps = &wpc->stream;
    int termcnt = wpmd->byte_length;
    int16_t wa = 0, wb = 0; /* local dpp */
    int counter = 0, dpp_idx, myiterator;
    if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) termcnt /= 2;
    if (termcnt > wps->num_terms) return 0;
    myiterator = wps->num_terms;
    while (termcnt > 0) {
        dpp_idx = myiterator - 1;
        wa = (int16_t)restore_weight((int8_t)ba_get(wpc, wpmd->data, counter));
        if (dpp_idx < 0 || dpp_idx >= MAX_NTERMS) THROW(wpc);
        wps->decorr_passes[dpp_idx].weight_A = wa;
        counter++;
        if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) {
            wb = (int16_t)restore_weight((int8_t)ba_get(wpc, wpmd->data, counter));
            counter++;
        }
        wps->decorr_passes[dpp_idx].weight_B = wb;
        myiterator--;
        termcnt--;
    }
    return 1;
}

static i32 rd16s(rd_context *c, ByteArray *a, int at) /* exp2s((short)(b0 + (b1<<8))) */
{
    int b0 = ba_get(c, a, at), b1 = ba_get(c, a, at + 1);
    return exp2s((int16_t)(b0 + (b1 << 8)));
}

static int read_decorr_samples(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:250-360 */
{
    WavpackStream *wps = &wpc->stream;
    ByteArray *byteptr = wpmd->data;
    decorr_pass dpp; /* local; .term is set by the zeroing loop and never refreshed (quirk C-1) */
    int tcount, counter = 0, dpp_index = 0;
    memset(&dpp, 0, sizeof(dpp));
    for (tcount = wps->num_terms; tcount > 0; tcount--) {
        dpp.term = wps->decorr_passes[dpp_index].term;
        memset(dpp.samples_A, 0, sizeof(dpp.samples_A));
        memset(dpp.samples_B, 0, sizeof(dpp.samples_B));
        memset(wps->decorr_passes[dpp_index].samples_A, 0, sizeof(dpp.samples_A));
        memset(wps->decorr_passes[dpp_index].samples_B, 0, sizeof(dpp.samples_B));
        dpp_index++;
    }
    if (wps->wphdr.version == 0x402 && (wps->wphdr.flags & HYBRID_FLAG) > 0) {
        counter += 2;
        if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) counter += 2;
    }
    dpp_index--;
    while (counter < wpmd->byte_length) {
        if (dpp.term > MAX_TERM) {
            i32 a0 = rd16s(wpc, byteptr, counter); /* operands are read before assignment in C# too */
            ba_get(wpc, byteptr, counter + 3);
            dpp.samples_A[0] = a0;
            dpp.samples_A[1] = rd16s(wpc, byteptr, counter + 2);
            counter += 4;
            if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) {
                ba_get(wpc, byteptr, counter + 3);
                dpp.samples_B[0] = rd16s(wpc, byteptr, counter);
                dpp.samples_B[1] = rd16s(wpc, byteptr, counter + 2);
                counter += 4;
            }
        } else if (dpp.term < 0) {
            ba_get(wpc, byteptr, counter + 3);
            dpp.samples_A[0] = rd16s(wpc, byteptr, counter);
            dpp.samples_B[0] = rd16s(wpc, byteptr, counter + 2);
            counter += 4;
        } else {
            int m = 0, cnt = dpp.term;
            while (cnt > 0) {
                dpp.samples_A[m] = rd16s(wpc, byteptr, counter);
                counter += 2;
                if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) {
                    dpp.samples_B[m] = rd16s(wpc, byteptr, counter);
                    counter += 2;
                }
                m++;
                cnt--;
            }
        }
        if (dpp_index < 0 || dpp_index >= MAX_NTERMS) THROW(wpc);
        memcpy(wps->decorr_passes[dpp_index].samples_A, dpp.samples_A, sizeof(dpp.samples_A));
        memcpy(wps->decorr_passes[dpp_index].samples_B, dpp.samples_B, sizeof(dpp.samples_B));
        dpp_index--;
    }
    return 1;
}

static int read_int32_info(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:367-382 */
{
    WavpackStream *wps = &wpc->stream;
    if (wpmd->byte_length != 4) return 0;
    wps->int32_sent_bits = ba_get(wpc, wpmd->data, 0);
    wps->int32_zeros = ba_get(wpc, wpmd->data, 1);
    wps->int32_ones = ba_get(wpc, wpmd->data, 2);
    wps->int32_dups = ba_get(wpc, wpmd->data, 3);
    return 1;
}

static int read_channel_info(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:389-410 */
{
    int bytecnt = wpmd->byte_length, shift = 0, counter = 0;
    i64 mask = 0;
    if (bytecnt == 0 || bytecnt > 5) return 0;
    wpc->config.num_channels = ba_get(wpc, wpmd->data, counter++);
    while (bytecnt >= 0) { /* over-reads two bytes of the shared buffer (quirk C-11) */
        mask |= (i64)shl32((i32)ba_get(wpc, wpmd->data, counter++), shift);
        shift += 8;
        bytecnt--;
    }
    wpc->config.channel_mask = mask;
    return 1;
}

static int read_new_config_info(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:415-427 */
{
    wpc->five = 1;
    if (wpmd->byte_length >= 1) wpc->file_format = ba_get(wpc, wpmd->data, 0);
    return 1;
}

static int read_config_info(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:432-455 */
{
    int bytecnt = wpmd->byte_length, counter = 0;
    if (bytecnt >= 3) {
        wpc->config.flags &= 0xff;
        wpc->config.flags |= (i64)(ba_get(wpc, wpmd->data, counter++) << 8);
        wpc->config.flags |= (i64)(ba_get(wpc, wpmd->data, counter++) << 16);
        wpc->config.flags |= (i64)shl32((i32)ba_get(wpc, wpmd->data, counter++), 24); /* int << 24 then sign-extended */
    }
    if (bytecnt >= 4 && (wpc->config.flags & CONFIG_EXTRA_MODE) > 0) {
        wpc->config.xmode = ba_get(wpc, wpmd->data, counter++);
        bytecnt--;
    }
    if (bytecnt >= 5) wpc->five = 1;
    return 1;
}

static int read_sample_rate(rd_context *wpc, WavpackMetadata *wpmd) /* UnpackUtils.cs:459-473 */
{
    if (wpmd->byte_length == 3) {
        wpc->config.sample_rate = ba_get(wpc, wpmd->data, 0);
        wpc->config.sample_rate |= (i64)(ba_get(wpc, wpmd->data, 1) << 8);
        wpc->config.sample_rate |= (i64)(ba_get(wpc, wpmd->data, 2) << 16);
    }
    return 1;
}

static void copy_bytes(rd_context *wpc, WavpackMetadata *wpmd, uint8_t **dst, long *dlen) /* UnpackUtils.cs:475-491 */
{
    if (wpmd->byte_length < 0) THROW(wpc); /* new byte[-1] */
    if (!wpmd->data || wpmd->byte_length > wpmd->data->len) THROW(wpc);
    free(*dst);
    *dst = (uint8_t *)malloc((size_t)wpmd->byte_length + 1);
    memcpy(*dst, wpmd->data->p, (size_t)wpmd->byte_length);
    *dlen = wpmd->byte_length;
}

static int read_float_info(rd_context *wpc, WavpackMetadata *wpmd) /* FloatUtils.cs:15-30 */
{
    WavpackStream *wps = &wpc->stream;
    if (wpmd->byte_length != 4) return 0;
    wps->float_flags = ba_get(wpc, wpmd->data, 0);
    wps->float_shift = ba_get(wpc, wpmd->data, 1);
    wps->float_max_exp = ba_get(wpc, wpmd->data, 2);
    wps->float_norm_exp = ba_get(wpc, wpmd->data, 3);
    return 1;
}

static i32 rd16u(rd_context *c, ByteArray *a, int at) { return ba_get(c, a, at) + (ba_get(c, a, at + 1) << 8); }

static int read_entropy_vars(rd_context *wpc, WavpackMetadata *wpmd) /* WordsUtils.cs:75-116 */
{
    WavpackStream *wps = &wpc->stream;
    int b[12];
    words_data w;
    memset(&w, 0, sizeof(w));
    for (int i = 0; i < 6; i++) b[i] = ba_get(wpc, wpmd->data, i);
    if (wpmd->byte_length != 12)
        if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0)
            return 0;
    w.c[0].median[0] = exp2s(b[0] + (b[1] << 8));
    w.c[0].median[1] = exp2s(b[2] + (b[3] << 8));
    w.c[0].median[2] = exp2s(b[4] + (b[5] << 8));
    if ((wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0) {
        for (int i = 6; i < 12; i++) b[i] = ba_get(wpc, wpmd->data, i);
        w.c[1].median[0] = exp2s(b[6] + (b[7] << 8));
        w.c[1].median[1] = exp2s(b[8] + (b[9] << 8));
        w.c[1].median[2] = exp2s(b[10] + (b[11] << 8));
    }
    wps->w = w;
    return 1;
}

static int read_hybrid_profile(rd_context *wpc, WavpackMetadata *wpmd) /* WordsUtils.cs:124-187 */
{
    WavpackStream *wps = &wpc->stream;
    ByteArray *p = wpmd->data;
    int bytecnt = wpmd->byte_length, k = 0;
    int stereo = (wps->wphdr.flags & (MONO_FLAG | FALSE_STEREO)) == 0;
    if (wps->wphdr.flags & HYBRID_BITRATE) {
        wps->w.c[0].slow_level = exp2s(rd16u(wpc, p, k)); k += 2;
        if (stereo) { wps->w.c[1].slow_level = exp2s(rd16u(wpc, p, k)); k += 2; }
    }
    wps->w.bitrate_acc[0] = (i64)shl32(rd16u(wpc, p, k), 16); k += 2; /* int << 16, then widened */
    if (stereo) { wps->w.bitrate_acc[1] = (i64)shl32(rd16u(wpc, p, k), 16); k += 2; }
    if (k < bytecnt) {
        wps->w.bitrate_delta[0] = exp2s((int16_t)rd16u(wpc, p, k)); k += 2;
        if (stereo) { wps->w.bitrate_delta[1] = exp2s((int16_t)rd16u(wpc, p, k)); k += 2; }
        if (k < bytecnt) return 0;
    } else
        wps->w.bitrate_delta[0] = wps->w.bitrate_delta[1] = 0;
    return 1;
}

/* ------------------------------------------------------------------ */
/* DsdUtils.cs init                                                    */
/* ------------------------------------------------------------------ */
enum { MAX_HISTORY_BITS = 5, MAX_BYTES_PER_BIN = 1280, MAX_DSD_BITS_VALUE = 256 };
enum { PTABLE_BITS = 8, PTABLE_BINS = 1 << PTABLE_BITS, PTABLE_MASK = PTABLE_BINS - 1, DSD_UP = 0x010000FE,
       DSD_DOWN = 0x00010000, DSD_DECAY = 8, DSD_PRECISION = 20, DSD_VALUE_ONE = 1 << DSD_PRECISION,
       DSD_PRECISION_USE = 12, DSD_RATE_S = 20 };

static void dsd_free(dsds *d)
{
    ba_unref(d->data);
    free(d->probabilities); free(d->lookup_buffer); free(d->value_lookup); free(d->summed_probabilities); free(d->ptable);
    memset(d, 0, sizeof(*d));
}

static int init_dsd_block_fast(rd_context *wpc) /* DsdUtils.cs:149-242 */
{
    dsds *d = &wpc->stream.dsd;
    int total_summed_probabilities = 0, bi, i;
    uint8_t max_probability;
    if (d->byteptr == d->data->len) return 0;
    uint8_t history_bits = ba_get(wpc, d->data, d->byteptr++);
    if (d->byteptr == d->data->len || history_bits > MAX_HISTORY_BITS) return 0;
    d->history_bins = 1 << history_bits;
    d->lookup_len = d->history_bins * MAX_BYTES_PER_BIN;
    d->lookup_buffer = (uint8_t *)calloc((size_t)d->lookup_len, 1);
    d->value_lookup = (i32 *)calloc((size_t)d->history_bins, sizeof(i32));
    d->summed_probabilities = (uint16_t *)calloc((size_t)MAX_DSD_BITS_VALUE * d->history_bins, sizeof(uint16_t));
    d->probabilities_len = MAX_DSD_BITS_VALUE * d->history_bins;
    d->probabilities = (uint8_t *)calloc((size_t)d->probabilities_len, 1);
    max_probability = ba_get(wpc, d->data, d->byteptr++);
    if (max_probability < 0xFF) {
        int outptr = 0, outend = d->probabilities_len;
        while (outptr < outend && d->byteptr < d->data->len) {
            uint8_t code = ba_get(wpc, d->data, d->byteptr++);
            if (code > max_probability) {
                int zcount = code - max_probability;
                while (outptr < outend && zcount-- > 0) d->probabilities[outptr++] = 0;
            } else if (code != 0)
                d->probabilities[outptr++] = code;
            else
                break;
        }
        if (outptr < outend || (d->byteptr < d->data->len && ba_get(wpc, d->data, d->byteptr++) > 0))
            return 0;
    } else if (d->data->len - d->byteptr > d->probabilities_len) {
        memcpy(d->probabilities, d->data->p + d->byteptr, (size_t)d->probabilities_len);
        d->byteptr += d->probabilities_len;
    } else
        return 0;

    int lb_ptr = 0;
    for (bi = 0; bi < d->history_bins; ++bi) {
        uint16_t sum_values;
        int bi_index = bi * MAX_DSD_BITS_VALUE;
        for (sum_values = 0, i = 0; i < MAX_DSD_BITS_VALUE; ++i)
            d->summed_probabilities[bi_index + i] = sum_values = (uint16_t)(sum_values + d->probabilities[bi_index + i]);
        if (sum_values != 0) {
            if ((total_summed_probabilities += sum_values) > d->history_bins * MAX_BYTES_PER_BIN) return 0;
            d->value_lookup[bi] = lb_ptr;
            for (i = 0; i < MAX_DSD_BITS_VALUE; i++) {
                int cc = d->probabilities[bi_index + i];
                while (cc-- > 0) {
                    if (lb_ptr >= d->lookup_len) THROW(wpc);
                    d->lookup_buffer[lb_ptr++] = (uint8_t)i;
                }
            }
        }
    }
    if (d->data->len - d->byteptr < 4 || total_summed_probabilities > d->history_bins * MAX_BYTES_PER_BIN) return 0;
    for (i = 4; i > 0; i--) d->value = (d->value << 8) | ba_get(wpc, d->data, d->byteptr++);
    d->p0 = d->p1 = 0;
    d->low = 0;
    d->high = 0xFFFFFFFFu;
    d->ready = 1;
    return 1;
}

static void init_ptable(i32 *table, int rate_i, int rate_s) /* DsdUtils.cs:321-341 */
{
    i32 value = 0x808000, rate = rate_i << 8, c, i;
    for (c = (rate + 128) >> 8; c > 0; c--) value += (DSD_DOWN - value) >> DSD_DECAY;
    for (i = 0; i < PTABLE_BINS / 2; ++i) {
        table[i] = value;
        table[PTABLE_BINS - 1 - i] = 0x100ffff - value;
        if (value > 0x010000) {
            rate += (rate * rate_s + 128) >> 8;
            for (c = (rate + 64) >> 7; c > 0; c--) value += (DSD_DOWN - value) >> DSD_DECAY;
        }
    }
}

static int init_dsd_block_high(rd_context *wpc) /* DsdUtils.cs:343-389 */
{
    WavpackStream *wps = &wpc->stream;
    dsds *d = &wps->dsd;
    u32 flags = wps->wphdr.flags;
    int channel, rate_i, rate_s, i;
    if (d->data->len - d->byteptr < ((flags & MONO_DATA) > 0 ? 13 : 20)) return 0;
    rate_i = ba_get(wpc, d->data, d->byteptr++);
    rate_s = ba_get(wpc, d->data, d->byteptr++);
    if (rate_s != DSD_RATE_S) return 0;
    if (!d->ptable) d->ptable = (i32 *)calloc(PTABLE_BINS, sizeof(i32));
    if (!d->filters_present) { memset(d->filters, 0, sizeof(d->filters)); d->filters_present = 1; }
    init_ptable(d->ptable, rate_i, rate_s);
    for (channel = 0; channel < ((flags & MONO_DATA) > 0 ? 1 : 2); ++channel) {
        DSDfilters *sp = &d->filters[channel];
        sp->filter1 = ba_get(wpc, d->data, d->byteptr++) << (DSD_PRECISION - 8);
        sp->filter2 = ba_get(wpc, d->data, d->byteptr++) << (DSD_PRECISION - 8);
        sp->filter3 = ba_get(wpc, d->data, d->byteptr++) << (DSD_PRECISION - 8);
        sp->filter4 = ba_get(wpc, d->data, d->byteptr++) << (DSD_PRECISION - 8);
        sp->filter5 = ba_get(wpc, d->data, d->byteptr++) << (DSD_PRECISION - 8);
        sp->filter6 = 0;
        sp->factor = ba_get(wpc, d->data, d->byteptr++);
        sp->factor |= ba_get(wpc, d->data, d->byteptr++) << 8;
        sp->factor = (i32)((u32)sp->factor << 16) >> 16;
    }
    d->high = 0xFFFFFFFFu;
    d->low = 0;
    for (i = 4; i > 0; i--) d->value = (d->value << 8) | ba_get(wpc, d->data, d->byteptr++);
    d->ready = 1;
    return 1;
}

static int init_dsd_block(rd_context *wpc, WavpackMetadata *wpmd) /* DsdUtils.cs:17-54 */
{
    WavpackStream *wps = &wpc->stream;
    if (wpmd->byte_length < 2 || ba_get(wpc, wpmd->data, 0) > 31) return 0;
    if (!copy_data(wpmd)) return 0;
    dsd_free(&wps->dsd); /* wps.dsd = new dsds() { data = wpmd.data } */
    wps->dsd.data = wpmd->data;
    ba_ref(wpmd->data);
    wpc->dsd_multiplier = 1U << (ba_get(wpc, wps->dsd.data, wps->dsd.byteptr++) & 31);
    wps->dsd.mode = ba_get(wpc, wps->dsd.data, wps->dsd.byteptr++);
    if (wps->dsd.mode == 0) {
        if ((i64)(wps->dsd.data->len - wps->dsd.byteptr) != (i64)wps->wphdr.block_samples * ((wps->wphdr.flags & MONO_DATA) > 0 ? 1 : 2))
            return 0;
        wps->dsd.ready = 1;
        return 1;
    } else if (wps->dsd.mode == 1)
        return init_dsd_block_fast(wpc);
    else if (wps->dsd.mode == 3)
        return init_dsd_block_high(wpc);
    return 0;
}

static int process_metadata(rd_context *wpc, WavpackMetadata *wpmd) /* MetadataUtils.cs:111-193 */
{
    switch (wpmd->id) {
    case ID_DUMMY: return 1;
    case ID_DECORR_TERMS: return read_decorr_terms(wpc, wpmd);
    case ID_DECORR_WEIGHTS: return read_decorr_weights(wpc, wpmd);
    case ID_DECORR_SAMPLES: return read_decorr_samples(wpc, wpmd);
    case ID_ENTROPY_VARS: return read_entropy_vars(wpc, wpmd);
    case ID_HYBRID_PROFILE: return read_hybrid_profile(wpc, wpmd);
    case ID_SHAPING_WEIGHTS: return 1;
    case ID_FLOAT_INFO: return read_float_info(wpc, wpmd);
    case ID_INT32_INFO: return read_int32_info(wpc, wpmd);
    case ID_CHANNEL_INFO: return read_channel_info(wpc, wpmd);
    case ID_CONFIG_BLOCK: return read_config_info(wpc, wpmd);
    case ID_SAMPLE_RATE: return read_sample_rate(wpc, wpmd);
    case ID_WV_BITSTREAM: return init_wv_bitstream(wpc, wpmd);
    case ID_WVC_BITSTREAM: return init_wvc_bitstream(wpc, wpmd);
    case ID_WVX_BITSTREAM:
    case ID_WVX_NEW_BITSTREAM: return init_wvx_bitstream(wpc, wpmd);
    case ID_DSD_BLOCK: return init_dsd_block(wpc, wpmd);
    case ID_NEW_CONFIG_BLOCK: return read_new_config_info(wpc, wpmd);
    case ID_RIFF_HEADER:
    case ID_ALT_HEADER: copy_bytes(wpc, wpmd, &wpc->header, &wpc->header_len); return 1;
    case ID_RIFF_TRAILER:
    case ID_ALT_TRAILER: copy_bytes(wpc, wpmd, &wpc->trailer, &wpc->trailer_len); return 1;
    case ID_ALT_EXTENSION: {
        if (wpmd->byte_length < 0 || !wpmd->data || wpmd->byte_length > wpmd->data->len) THROW(wpc);
        free(wpc->file_extension);
        wpc->file_extension = (char *)calloc((size_t)wpmd->byte_length + 1, 1);
        memcpy(wpc->file_extension, wpmd->data->p, (size_t)wpmd->byte_length);
        return 1;
    }
    case ID_BLOCK_CHECKSUM: wpc->five = 1; return 1;
    default: return (wpmd->id & ID_OPTIONAL_DATA) != 0;
    }
}

/* ------------------------------------------------------------------ */
/* UnpackUtils.unpack_init                                             */
/* ------------------------------------------------------------------ */
static int unpack_init(rd_context *wpc) /* UnpackUtils.cs:24-68 */
{
    WavpackStream *wps = &wpc->stream;
    WavpackMetadata *wpmd = &wpc->md; /* `new WavpackMetadata()`; lives in the context so an "exception" can release it */
    md_set_data(wpmd, NULL);
    memset(wpmd, 0, sizeof(*wpmd));
    wpmd->bytecount = 24;

    if (wps->wphdr.block_samples > 0 && wps->wphdr.block_index != 0xFFFFFFFFLL)
        wps->sample_index = wps->wphdr.block_index;
    wps->mute_error = 0;
    wps->crc = wps->crc_x = -1;
    wps->wvbits.sr = 0;

    while (read_metadata_buff(wpc, wpmd)) {
        if (!process_metadata(wpc, wpmd)) {
            snprintf(wpc->error_buf, sizeof(wpc->error_buf), "invalid metadata id %d", wpmd->id);
            wpc->error_message = wpc->error_buf;
            md_set_data(wpmd, NULL);
            return 0;
        }
    }
    i64 bytecount = wpmd->bytecount;
    md_set_data(wpmd, NULL);

    if (bytecount != (i64)wps->wphdr.ckSize) {
        wpc->error_message = "invalid reading WavPack metadata block";
        return 0;
    }
    if ((wps->wphdr.block_samples != 0 && (wps->wphdr.flags & DSD_FLAG) > 0) ? !wps->dsd.ready
                                                                            : (wps->wvbits.is_null || wps->wvbits.end == 0)) {
        wpc->error_message = "invalid WavPack file";
        return 0;
    }
    if (wps->wphdr.block_samples != 0) {
        if ((wps->wphdr.flags & INT32_DATA) != 0 && wps->int32_sent_bits != 0 && wps->wvxbits.is_null)
            wpc->lossy_blocks = 1;
        if ((wps->wphdr.flags & FLOAT_DATA) != 0 &&
            (wps->float_flags & (FLOAT_EXCEPTIONS | FLOAT_ZEROS_SENT | FLOAT_SHIFT_SENT | FLOAT_SHIFT_SAME)) != 0)
            wpc->lossy_blocks = 1;
    }
    return 1;
}

/* ------------------------------------------------------------------ */
/* WordsUtils: update_error_limit / read_code / get_words              */
/* ------------------------------------------------------------------ */
enum { LIMIT_ONES = 16, SLS = 8, SLO = 1 << (SLS - 1), DIV0 = 128, DIV1 = 64, DIV2 = 32 };

static void update_error_limit(words_data *w, i64 flags) /* WordsUtils.cs:195-261 */
{
    i32 bitrate_0 = (i32)((w->bitrate_acc[0] += w->bitrate_delta[0]) >> 16);
    if ((flags & (MONO_FLAG | FALSE_STEREO)) != 0) {
        if ((flags & HYBRID_BITRATE) != 0) {
            i32 slow_log_0 = (w->c[0].slow_level + SLO) >> SLS;
            if (slow_log_0 - bitrate_0 > -0x100)
                w->c[0].error_limit = exp2s(slow_log_0 - bitrate_0 + 0x100);
            else
                w->c[0].error_limit = 0;
        } else
            w->c[0].error_limit = exp2s(bitrate_0);
    } else {
        i32 bitrate_1 = (i32)((w->bitrate_acc[1] += w->bitrate_delta[1]) >> 16);
        if ((flags & HYBRID_BITRATE) != 0) {
            i32 slow_log_0 = (w->c[0].slow_level + SLO) >> SLS;
            i32 slow_log_1 = (w->c[1].slow_level + SLO) >> SLS;
            if ((flags & HYBRID_BALANCE) != 0) {
                i32 balance = (slow_log_1 - slow_log_0 + bitrate_1 + 1) >> 1;
                if (balance > bitrate_0) {
                    bitrate_1 = bitrate_0 * 2;
                    bitrate_0 = 0;
                } else if (-balance > bitrate_0) {
                    bitrate_0 = bitrate_0 * 2;
                    bitrate_1 = 0;
                } else {
                    bitrate_1 = bitrate_0 + balance;
                    bitrate_0 = bitrate_0 - balance;
                }
            }
            if (slow_log_0 - bitrate_0 > -0x100)
                w->c[0].error_limit = exp2s(slow_log_0 - bitrate_0 + 0x100);
            else
                w->c[0].error_limit = 0;
            if (slow_log_1 - bitrate_1 > -0x100)
                w->c[1].error_limit = exp2s(slow_log_1 - bitrate_1 + 0x100);
            else
                w->c[1].error_limit = 0;
        } else {
            w->c[0].error_limit = exp2s(bitrate_0);
            w->c[1].error_limit = exp2s(bitrate_1);
        }
    }
}

static i64 read_code(rd_context *c, Bitstream *bs, i64 maxcode) /* WordsUtils.cs:546-570 */
{
    int bitcount = count_bits(c, maxcode);
    i64 extras = (i64)shl32(1, bitcount) - maxcode - 1;
    i64 code;
    if (bitcount == 0) return 0;
    code = getbits(c, bitcount - 1, bs);
    code &= (i64)(i32)((u32)shl32(1, bitcount - 1) - 1u);
    if (code >= extras) {
        code = (code << 1) - extras;
        if (getbit(c, bs)) ++code;
    }
    return code;
}

static int get_words(rd_context *wpc, i64 nsamples, i64 flags, words_data *w, Bitstream *bs, i32 *buffer, long buffer_len,
                     int bufferStartPos) /* WordsUtils.cs:272-511 */
{
    entropy_data *c = w->c;
    int csamples;
    int buffer_counter = bufferStartPos;
    int entidx = 1;
    const int mono = (flags & (MONO_FLAG | FALSE_STEREO)) != 0;

    if (!mono) nsamples *= 2; else entidx = 0;

#define BUF(i) (*(((i) < 0 || (i) >= buffer_len) ? (THROW(wpc), (i32 *)0) : &buffer[(i)]))

    for (csamples = 0; csamples < nsamples; ++csamples) {
        int ones_count;
        i64 low, high, mid;

        if (!mono) entidx = (entidx == 1) ? 0 : 1;

        if ((w->c[0].median[0] & ~1) == 0 && !w->holding_zero && !w->holding_one && (w->c[1].median[0] & ~1) == 0) {
            i64 mask;
            int cbits;
            if (w->zeros_acc > 0) {
                if (--w->zeros_acc > 0) {
                    c[entidx].slow_level -= (c[entidx].slow_level + SLO) >> SLS;
                    BUF(buffer_counter) = 0;
                    buffer_counter++;
                    continue;
                }
            } else {
                for (cbits = 0; cbits < 33 && getbit(wpc, bs); ++cbits);
                if (cbits == 33) break;
                if (cbits < 2)
                    w->zeros_acc = cbits;
                else {
                    for (mask = 1, w->zeros_acc = 0; --cbits > 0; mask <<= 1)
                        if (getbit(wpc, bs)) w->zeros_acc |= mask;
                    w->zeros_acc |= mask;
                }
                if (w->zeros_acc > 0) {
                    c[entidx].slow_level -= ((c[entidx].slow_level + SLO) >> SLS);
                    w->c[0].median[0] = w->c[0].median[1] = w->c[0].median[2] = 0;
                    w->c[1].median[0] = w->c[1].median[1] = w->c[1].median[2] = 0;
                    BUF(buffer_counter) = 0;
                    buffer_counter++;
                    continue;
                }
            }
        }

        if (w->holding_zero) {
            w->holding_zero = 0;
            ones_count = 0;
        } else {
            if (bs->bc < 8) {
                bs->ptr++;
                bs->buf_index++;
                if (bs->ptr == bs->end) bs_read(wpc, bs);
                bs->sr |= (u32)shl32((i32)ba_get(wpc, bs->buf, bs->buf_index), bs->bc);
                bs->bc += 8;
            }
            uint8_t next8 = (uint8_t)bs->sr;
            if (next8 == 0xff) {
                bs->bc -= 8;
                bs->sr >>= 8;
                for (ones_count = 8; ones_count < (LIMIT_ONES + 1) && getbit(wpc, bs); ++ones_count);
                if (ones_count == (LIMIT_ONES + 1)) break;
                if (ones_count == LIMIT_ONES) {
                    int mask, cbits;
                    for (cbits = 0; cbits < 33 && getbit(wpc, bs); ++cbits);
                    if (cbits == 33) break;
                    if (cbits < 2)
                        ones_count = cbits;
                    else {
                        for (mask = 1, ones_count = 0; --cbits > 0; mask = shl32(mask, 1))
                            if (getbit(wpc, bs)) ones_count |= mask;
                        ones_count |= mask;
                    }
                    ones_count += LIMIT_ONES;
                }
            } else {
                bs->bc -= (ones_count = ones_count_table[next8]) + 1;
                bs->sr = shru32(bs->sr, ones_count + 1);
            }
            if (w->holding_one) {
                w->holding_one = (ones_count & 1) > 0;
                ones_count = (ones_count >> 1) + 1;
            } else {
                w->holding_one = (ones_count & 1) > 0;
                ones_count >>= 1;
            }
            w->holding_zero = !w->holding_one;
        }

        if ((flags & HYBRID_FLAG) > 0 && (mono || (csamples & 1) == 0))
            update_error_limit(w, flags);

        if (ones_count == 0) {
            low = 0;
            high = (((c[entidx].median[0]) >> 4) + 1) - 1;
            c[entidx].median[0] -= (((c[entidx].median[0] + (DIV0 - 2)) >> 7) * 2);
        } else {
            low = (((c[entidx].median[0]) >> 4) + 1);
            c[entidx].median[0] += ((c[entidx].median[0] + DIV0) >> 7) * 5;
            if (ones_count == 1) {
                high = low + (((c[entidx].median[1]) >> 4) + 1) - 1;
                c[entidx].median[1] -= ((c[entidx].median[1] + (DIV1 - 2)) >> 6) * 2;
            } else {
                low += (((c[entidx].median[1]) >> 4) + 1);
                c[entidx].median[1] += ((c[entidx].median[1] + DIV1) >> 6) * 5;
                if (ones_count == 2) {
                    high = low + (((c[entidx].median[2]) >> 4) + 1) - 1;
                    c[entidx].median[2] -= ((c[entidx].median[2] + (DIV2 - 2)) >> 5) * 2;
                } else {
                    low += (i32)((ones_count - 2) * (((c[entidx].median[2]) >> 4) + 1)); /* int product */
                    high = low + (((c[entidx].median[2]) >> 4) + 1) - 1;
                    c[entidx].median[2] += ((c[entidx].median[2] + DIV2) >> 5) * 5;
                }
            }
        }

        mid = (high + low + 1) >> 1;

        if (c[entidx].error_limit == 0) {
            mid = read_code(wpc, bs, high - low);
            mid = mid + low;
        } else
            while (high - low > c[entidx].error_limit) {
                if (getbit(wpc, bs))
                    mid = (high + (low = mid) + 1) >> 1;
                else
                    mid = ((high = mid - 1) + low + 1) >> 1;
            }

        if (getbit(wpc, bs))
            BUF(buffer_counter) = (i32)~mid;
        else
            BUF(buffer_counter) = (i32)mid;
        buffer_counter++;

        if ((flags & HYBRID_BITRATE) > 0)
            c[entidx].slow_level = c[entidx].slow_level - ((c[entidx].slow_level + SLO) >> SLS) + mylog2(wpc, mid);
    }
#undef BUF
    if (mono) return csamples;
    return csamples / 2;
}

/* ------------------------------------------------------------------ */
/* Decorrelation passes                                                */
/* ------------------------------------------------------------------ */
#define APPLY_W(w, s) ((i32)(((i64)(w) * (i64)(s) + 512) >> 10))

static inline void upd_w(int *w, int delta, i32 sam, i32 in) /* `if (sam != 0 && in != 0) w += sign * delta` */
{
    if (sam != 0 && in != 0) {
        if ((sam ^ in) < 0) *w -= delta; else *w += delta;
    }
}
static inline void upd_w_clip(int *w, int delta, i32 sam, i32 in) /* UnpackUtils.cs:776-785 pattern */
{
    if ((sam ^ in) < 0) {
        if (sam != 0 && in != 0 && (*w -= delta) < -1024) *w = (*w < 0) ? -1024 : 1024;
    } else {
        if (sam != 0 && in != 0 && (*w += delta) > 1024) *w = (*w < 0) ? -1024 : 1024;
    }
}

#define CHK(lo, hi) do { if ((lo) < 0 || (hi) > buffer_len) THROW(wpc); } while (0)

static void decorr_stereo_pass(rd_context *wpc, decorr_pass *dpp, i32 *buffer, long buffer_len, i64 sample_count, int buf_idx)
/* UnpackUtils.cs:688-944 */
{
    int delta = dpp->delta, weight_A = dpp->weight_A, weight_B = dpp->weight_B;
    i32 sam_A, sam_B;
    int m, k;
    i64 b, end = buf_idx + sample_count * 2;
    if (sample_count > 0) CHK(buf_idx, end);

    switch (dpp->term) {
    case 17:
    case 18:
        for (b = buf_idx; b < end; b += 2) {
            sam_A = dpp->term == 17 ? 2 * dpp->samples_A[0] - dpp->samples_A[1] : (3 * dpp->samples_A[0] - dpp->samples_A[1]) >> 1;
            dpp->samples_A[1] = dpp->samples_A[0];
            dpp->samples_A[0] = APPLY_W(weight_A, sam_A) + buffer[b];
            upd_w(&weight_A, delta, sam_A, buffer[b]);
            buffer[b] = dpp->samples_A[0];
            sam_A = dpp->term == 17 ? 2 * dpp->samples_B[0] - dpp->samples_B[1] : (3 * dpp->samples_B[0] - dpp->samples_B[1]) >> 1;
            dpp->samples_B[1] = dpp->samples_B[0];
            dpp->samples_B[0] = APPLY_W(weight_B, sam_A) + buffer[b + 1];
            upd_w(&weight_B, delta, sam_A, buffer[b + 1]);
            buffer[b + 1] = dpp->samples_B[0];
        }
        break;
    case -1:
        for (b = buf_idx; b < end; b += 2) {
            sam_A = buffer[b] + APPLY_W(weight_A, dpp->samples_A[0]);
            upd_w_clip(&weight_A, delta, dpp->samples_A[0], buffer[b]);
            buffer[b] = sam_A;
            dpp->samples_A[0] = buffer[b + 1] + APPLY_W(weight_B, sam_A);
            upd_w_clip(&weight_B, delta, sam_A, buffer[b + 1]);
            buffer[b + 1] = dpp->samples_A[0];
        }
        break;
    case -2:
        for (b = buf_idx; b < end; b += 2) {
            sam_B = buffer[b + 1] + APPLY_W(weight_B, dpp->samples_B[0]);
            upd_w_clip(&weight_B, delta, dpp->samples_B[0], buffer[b + 1]);
            buffer[b + 1] = sam_B;
            dpp->samples_B[0] = buffer[b] + APPLY_W(weight_A, sam_B);
            upd_w_clip(&weight_A, delta, sam_B, buffer[b]);
            buffer[b] = dpp->samples_B[0];
        }
        break;
    case -3:
        for (b = buf_idx; b < end; b += 2) {
            sam_A = buffer[b] + APPLY_W(weight_A, dpp->samples_A[0]);
            upd_w_clip(&weight_A, delta, dpp->samples_A[0], buffer[b]);
            sam_B = buffer[b + 1] + APPLY_W(weight_B, dpp->samples_B[0]);
            upd_w_clip(&weight_B, delta, dpp->samples_B[0], buffer[b + 1]);
            buffer[b] = dpp->samples_B[0] = sam_A;
            buffer[b + 1] = dpp->samples_A[0] = sam_B;
        }
        break;
    default:
        for (m = 0, k = dpp->term & (MAX_TERM - 1), b = buf_idx; b < end; b += 2) {
            sam_A = dpp->samples_A[m];
            dpp->samples_A[k] = APPLY_W(weight_A, sam_A) + buffer[b];
            upd_w(&weight_A, delta, sam_A, buffer[b]);
            buffer[b] = dpp->samples_A[k];
            sam_A = dpp->samples_B[m];
            dpp->samples_B[k] = APPLY_W(weight_B, sam_A) + buffer[b + 1];
            upd_w(&weight_B, delta, sam_A, buffer[b + 1]);
            buffer[b + 1] = dpp->samples_B[k];
            m = (m + 1) & (MAX_TERM - 1);
            k = (k + 1) & (MAX_TERM - 1);
        }
        if (m != 0) {
            i32 t[MAX_TERM];
            memcpy(t, dpp->samples_A, sizeof(t));
            for (k = 0; k < MAX_TERM; k++, m++) dpp->samples_A[k] = t[m & (MAX_TERM - 1)];
            memcpy(t, dpp->samples_B, sizeof(t));
            for (k = 0; k < MAX_TERM; k++, m++) dpp->samples_B[k] = t[m & (MAX_TERM - 1)];
        }
        break;
    }
    dpp->weight_A = (int16_t)weight_A;
    dpp->weight_B = (int16_t)weight_B;
}

static void decorr_stereo_pass_cont(rd_context *wpc, decorr_pass *dpp, i32 *buffer, long buffer_len, i64 sample_count, int buf_idx)
/* UnpackUtils.cs:946-1154 */
{
    int delta = dpp->delta, weight_A = dpp->weight_A, weight_B = dpp->weight_B;
    i64 tptr, bi = buf_idx, end_index = buf_idx + sample_count * 2;
    i32 sam_A, sam_B;
    int k, i;
    CHK(buf_idx - 16, end_index);

    switch (dpp->term) {
    case 17:
    case 18:
        for (bi = buf_idx; bi < end_index; bi += 2) {
            sam_A = dpp->term == 17 ? 2 * buffer[bi - 2] - buffer[bi - 4] : (3 * buffer[bi - 2] - buffer[bi - 4]) >> 1;
            buffer[bi] = APPLY_W(weight_A, sam_A) + (sam_B = buffer[bi]);
            if (sam_A != 0 && sam_B != 0) weight_A += (((sam_A ^ sam_B) >> 30) | 1) * delta;
            sam_A = dpp->term == 17 ? 2 * buffer[bi - 1] - buffer[bi - 3] : (3 * buffer[bi - 1] - buffer[bi - 3]) >> 1;
            buffer[bi + 1] = APPLY_W(weight_B, sam_A) + (sam_B = buffer[bi + 1]);
            if (sam_A != 0 && sam_B != 0) weight_B += (((sam_A ^ sam_B) >> 30) | 1) * delta;
        }
        dpp->samples_B[0] = buffer[bi - 1];
        dpp->samples_A[0] = buffer[bi - 2];
        dpp->samples_B[1] = buffer[bi - 3];
        dpp->samples_A[1] = buffer[bi - 4];
        break;
    case -1:
        for (bi = buf_idx; bi < end_index; bi += 2) {
            i32 p = buffer[bi - 1];
            buffer[bi] = APPLY_W(weight_A, p) + (sam_A = buffer[bi]);
            upd_w_clip(&weight_A, delta, p, sam_A);
            p = buffer[bi];
            buffer[bi + 1] = APPLY_W(weight_B, p) + (sam_A = buffer[bi + 1]);
            upd_w_clip(&weight_B, delta, p, sam_A);
        }
        dpp->samples_A[0] = buffer[bi - 1];
        break;
    case -2:
        for (bi = buf_idx; bi < end_index; bi += 2) {
            i32 p = buffer[bi - 2];
            buffer[bi + 1] = APPLY_W(weight_B, p) + (sam_A = buffer[bi + 1]);
            upd_w_clip(&weight_B, delta, p, sam_A);
            p = buffer[bi + 1];
            buffer[bi] = APPLY_W(weight_A, p) + (sam_A = buffer[bi]);
            upd_w_clip(&weight_A, delta, p, sam_A);
        }
        dpp->samples_B[0] = buffer[bi - 2];
        break;
    case -3:
        for (bi = buf_idx; bi < end_index; bi += 2) {
            i32 p = buffer[bi - 1];
            buffer[bi] = APPLY_W(weight_A, p) + (sam_A = buffer[bi]);
            upd_w_clip(&weight_A, delta, p, sam_A);
            p = buffer[bi - 2];
            buffer[bi + 1] = APPLY_W(weight_B, p) + (sam_A = buffer[bi + 1]);
            upd_w_clip(&weight_B, delta, p, sam_A);
        }
        dpp->samples_A[0] = buffer[bi - 1];
        dpp->samples_B[0] = buffer[bi - 2];
        break;
    default:
        tptr = buf_idx - (dpp->term * 2);
        for (bi = buf_idx; bi < end_index; bi += 2) {
            buffer[bi] = APPLY_W(weight_A, buffer[tptr]) + (sam_A = buffer[bi]);
            if (buffer[tptr] != 0 && sam_A != 0) weight_A += (((buffer[tptr] ^ sam_A) >> 30) | 1) * delta;
            buffer[bi + 1] = APPLY_W(weight_B, buffer[tptr + 1]) + (sam_A = buffer[bi + 1]);
            if (buffer[tptr + 1] != 0 && sam_A != 0) weight_B += (((buffer[tptr + 1] ^ sam_A) >> 30) | 1) * delta;
            tptr += 2;
        }
        bi--;
        for (k = dpp->term - 1, i = 8; i > 0; k--) {
            i--;
            dpp->samples_B[k & (MAX_TERM - 1)] = buffer[bi];
            bi--;
            dpp->samples_A[k & (MAX_TERM - 1)] = buffer[bi];
            bi--;
        }
        break;
    }
    dpp->weight_A = (int16_t)weight_A;
    dpp->weight_B = (int16_t)weight_B;
}

static void decorr_mono_pass(rd_context *wpc, decorr_pass *dpp, i32 *buffer, long buffer_len, i64 sample_count, int buf_idx)
/* UnpackUtils.cs:1156-1240.  Negative terms fall into `default` with k = term & 7. */
{
    int delta = dpp->delta, weight_A = dpp->weight_A;
    i32 sam_A;
    int m, k;
    i64 b, end = buf_idx + sample_count;
    if (sample_count > 0) CHK(buf_idx, end);

    switch (dpp->term) {
    case 17:
    case 18:
        for (b = buf_idx; b < end; b++) {
            sam_A = dpp->term == 17 ? 2 * dpp->samples_A[0] - dpp->samples_A[1] : (3 * dpp->samples_A[0] - dpp->samples_A[1]) >> 1;
            dpp->samples_A[1] = dpp->samples_A[0];
            dpp->samples_A[0] = APPLY_W(weight_A, sam_A) + buffer[b];
            upd_w(&weight_A, delta, sam_A, buffer[b]);
            buffer[b] = dpp->samples_A[0];
        }
        break;
    default:
        for (m = 0, k = dpp->term & (MAX_TERM - 1), b = buf_idx; b < end; b++) {
            sam_A = dpp->samples_A[m];
            dpp->samples_A[k] = APPLY_W(weight_A, sam_A) + buffer[b];
            upd_w(&weight_A, delta, sam_A, buffer[b]);
            buffer[b] = dpp->samples_A[k];
            m = (m + 1) & (MAX_TERM - 1);
            k = (k + 1) & (MAX_TERM - 1);
        }
        if (m != 0) {
            i32 t[MAX_TERM];
            memcpy(t, dpp->samples_A, sizeof(t));
            for (k = 0; k < MAX_TERM; k++, m++) dpp->samples_A[k] = t[m & (MAX_TERM - 1)];
        }
        break;
    }
    dpp->weight_A = (int16_t)weight_A;
}

/* ------------------------------------------------------------------ */
/* FloatUtils.float_values / UnpackUtils.fixup_samples                 */
/* ------------------------------------------------------------------ */
static void float_values(rd_context *wpc, WavpackStream *wps, i32 *values, long buffer_len, i64 num_values, int bufferStartPos)
/* FloatUtils.cs:32-56 */
{
    int shift = wps->float_max_exp - wps->float_norm_exp + wps->float_shift;
    i64 vc = bufferStartPos;
    if (shift > 32) shift = 32; else if (shift < -32) shift = -32;
    if (num_values > 0 && (vc < 0 || vc + num_values > buffer_len)) THROW(wpc);
    while (num_values-- > 0) {
        if (shift > 0) values[vc] = shl32(values[vc], shift);
        else if (shift < 0) values[vc] = sar32(values[vc], -shift);
        if (values[vc] > 8388607) values[vc] = 8388607;
        else if (values[vc] < -8388608) values[vc] = -8388608;
        vc++;
    }
}

static void fixup_samples(rd_context *wpc, WavpackStream *wps, i32 *buffer, long buffer_len, i64 sample_count, int bufferStartPos)
/* UnpackUtils.cs:1251-1404 */
{
    i64 flags = wps->wphdr.flags;
    int lossy_flag = (flags & HYBRID_FLAG) > 0;
    int shift = (int)((flags & SHIFT_MASK) >> SHIFT_LSB);

    if ((flags & FLOAT_DATA) > 0) {
        float_values(wpc, wps, buffer, buffer_len, (flags & MONO_FLAG) > 0 ? sample_count : sample_count * 2, bufferStartPos);
        return;
    }

    if ((flags & INT32_DATA) > 0) {
        i64 count = (flags & MONO_FLAG) > 0 ? sample_count : sample_count * 2;
        int sent_bits = wps->int32_sent_bits, zeros = wps->int32_zeros;
        int ones = wps->int32_ones, dups = wps->int32_dups;
        u32 data, mask = shlu32(1U, sent_bits) - 1;
        i64 bc = bufferStartPos;
        if (count > 0 && (bc < 0 || bc + count > buffer_len)) THROW(wpc);

        if (!wps->wvxbits.is_null) {
            int max_width = wps->int32_max_width;
            i32 crc = wps->crc_x;
            while (count-- > 0) {
                if (sent_bits > 0) {
                    if (max_width > 0) {
                        i32 pvalue = buffer[bc] < 0 ? ~buffer[bc] : buffer[bc];
                        int width = count_bits(wpc, pvalue) + sent_bits;
                        int bits_to_read = sent_bits;
                        if (width <= max_width || (bits_to_read -= width - max_width) > 0) {
                            data = (u32)getbits(wpc, bits_to_read, &wps->wvxbits) & mask;
                            buffer[bc] = shl32((i32)((u32)shl32(buffer[bc], bits_to_read) | data), sent_bits - bits_to_read);
                        } else
                            buffer[bc] = shl32(buffer[bc], sent_bits);
                    } else {
                        data = (u32)(getbits(wpc, sent_bits, &wps->wvxbits) & mask);
                        buffer[bc] = (i32)(shlu32((u32)buffer[bc], sent_bits) | data);
                    }
                }
                if (zeros != 0)
                    buffer[bc] = shl32(buffer[bc], zeros);
                else if (ones != 0)
                    buffer[bc] = shl32(buffer[bc] + 1, ones) - 1;
                else if (dups != 0)
                    buffer[bc] = shl32(buffer[bc] + (buffer[bc] & 1), dups) - (buffer[bc] & 1);
                crc = crc * 9 + (buffer[bc] & 0xffff) * 3 + ((buffer[bc] >> 16) & 0xffff);
                bc++;
            }
            wps->crc_x = crc;
        } else if (sent_bits == 0 && (zeros + ones + dups) != 0) {
            while (lossy_flag && (flags & BYTES_STORED) == 3 && shift < 8) {
                if (zeros > 0) zeros--;
                else if (ones > 0) ones--;
                else if (dups > 0) dups--;
                else break;
                shift++;
            }
            while (count-- > 0) {
                if (zeros != 0)
                    buffer[bc] = shl32(buffer[bc], zeros);
                else if (ones != 0)
                    buffer[bc] = shl32(buffer[bc] + 1, ones) - 1;
                else if (dups != 0)
                    buffer[bc] = shl32(buffer[bc] + (buffer[bc] & 1), dups) - (buffer[bc] & 1);
                bc++;
            }
        } else
            shift += zeros + sent_bits + ones + dups;
    }

    shift &= 0x1f;

    if (lossy_flag) {
        i32 min_value, max_value, min_shifted, max_shifted;
        i64 bc = bufferStartPos;
        switch (flags & BYTES_STORED) {
        case 0:
            min_shifted = shl32(min_value = -128 >> shift, shift);
            max_shifted = shl32(max_value = 127 >> shift, shift);
            break;
        case 1:
            min_shifted = shl32(min_value = -32768 >> shift, shift);
            max_shifted = shl32(max_value = 32767 >> shift, shift);
            break;
        case 2:
            min_shifted = shl32(min_value = -8388608 >> shift, shift);
            max_shifted = shl32(max_value = 8388607 >> shift, shift);
            break;
        case 3:
        default:
            min_shifted = shl32(min_value = (i32)(0x80000000u >> shift), shift); /* unsigned shift (quirk C-7) */
            max_shifted = shl32(max_value = 0x7FFFFFFF >> shift, shift);
            break;
        }
        if ((flags & MONO_FLAG) == 0) sample_count *= 2;
        if (sample_count > 0 && (bc < 0 || bc + sample_count > buffer_len)) THROW(wpc);
        while (sample_count-- > 0) {
            if (buffer[bc] < min_value) buffer[bc] = min_shifted;
            else if (buffer[bc] > max_value) buffer[bc] = max_shifted;
            else buffer[bc] = shl32(buffer[bc], shift);
            bc++;
        }
    } else if (shift != 0) {
        i64 bc = bufferStartPos;
        if ((flags & MONO_FLAG) == 0) sample_count *= 2;
        if (sample_count > 0 && (bc < 0 || bc + sample_count > buffer_len)) THROW(wpc);
        while (sample_count-- > 0) { buffer[bc] = shl32(buffer[bc], shift); bc++; }
    }
}

/* ------------------------------------------------------------------ */
/* UnpackUtils.unpack_samples                                          */
/* ------------------------------------------------------------------ */
static i64 unpack_samples(rd_context *wpc, i32 *buffer, long buffer_len, i64 sample_count, int bufferStartPos)
/* UnpackUtils.cs:510-686 */
{
    WavpackStream *wps = &wpc->stream;
    i64 flags = wps->wphdr.flags;
    i64 i;
    i32 crc = wps->crc;
    i32 mute_limit = (i32)((1LL << (int)((flags & MAG_MASK) >> MAG_LSB)) + 2);
    int tcount;
    i64 buffer_counter = 0;

    if (wps->sample_index + sample_count > wps->wphdr.block_index + wps->wphdr.block_samples)
        sample_count = wps->wphdr.block_index + wps->wphdr.block_samples - wps->sample_index;

    if (wps->mute_error) {
        i64 tempc = (flags & MONO_FLAG) > 0 ? sample_count : 2 * sample_count;
        buffer_counter = bufferStartPos;
        if (tempc > 0 && (buffer_counter < 0 || buffer_counter + tempc > buffer_len)) THROW(wpc);
        while (tempc-- > 0) buffer[buffer_counter++] = 0;
        wps->sample_index += sample_count;
        return sample_count;
    }

    if ((flags & HYBRID_FLAG) > 0) mute_limit *= 2;

    if ((flags & (MONO_FLAG | FALSE_STEREO)) > 0) {
        int dpp_index = 0;
        i = get_words(wpc, sample_count, flags, &wps->w, &wps->wvbits, buffer, buffer_len, bufferStartPos);
        for (tcount = wps->num_terms; tcount > 0; tcount--, dpp_index++)
            decorr_mono_pass(wpc, &wps->decorr_passes[dpp_index], buffer, buffer_len, sample_count, bufferStartPos);
        int crclimit = (int)(sample_count + bufferStartPos);
        if (crclimit > bufferStartPos) CHK(bufferStartPos, crclimit);
        for (int q = bufferStartPos; q < crclimit; q++) {
            i32 bf_i = buffer[q];
            i32 bf_abs = bf_i < 0 ? -bf_i : bf_i;
            if (bf_abs > mute_limit) {
                i = q; /* buffer index, not sample index (quirk C-5) */
                break;
            }
            crc = crc * 3 + bf_i;
        }
    } else {
        i = get_words(wpc, sample_count, flags, &wps->w, &wps->wvbits, buffer, buffer_len, bufferStartPos);
        int dpp_index = 0;
        if (sample_count < 16) {
            for (tcount = wps->num_terms; tcount > 0; tcount--, dpp_index++)
                decorr_stereo_pass(wpc, &wps->decorr_passes[dpp_index], buffer, buffer_len, sample_count, bufferStartPos);
        } else {
            for (tcount = wps->num_terms; tcount > 0; tcount--, dpp_index++) {
                decorr_stereo_pass(wpc, &wps->decorr_passes[dpp_index], buffer, buffer_len, 8, bufferStartPos);
                decorr_stereo_pass_cont(wpc, &wps->decorr_passes[dpp_index], buffer, buffer_len, sample_count - 8, bufferStartPos + 16);
            }
        }
        if (sample_count > 0) CHK(bufferStartPos, bufferStartPos + sample_count * 2);
        i32 *bp = buffer + bufferStartPos;
        if ((flags & JOINT_STEREO) > 0) {
            for (buffer_counter = 0; buffer_counter < sample_count * 2; buffer_counter += 2) {
                /* L += (R -= (L >> 1)), C# left-to-right operand evaluation (App. E-9) */
                i32 oldL = bp[buffer_counter];
                bp[buffer_counter + 1] -= (oldL >> 1);
                bp[buffer_counter] = oldL + bp[buffer_counter + 1];
                i32 a = bp[buffer_counter] < 0 ? -bp[buffer_counter] : bp[buffer_counter];
                i32 b1 = bp[buffer_counter + 1] < 0 ? -bp[buffer_counter + 1] : bp[buffer_counter + 1];
                if (a > mute_limit || b1 > mute_limit) { i = buffer_counter / 2; break; }
                crc = (crc * 3 + bp[buffer_counter]) * 3 + bp[buffer_counter + 1];
            }
        } else {
            for (buffer_counter = 0; buffer_counter < sample_count * 2; buffer_counter += 2) {
                i32 a = bp[buffer_counter] < 0 ? -bp[buffer_counter] : bp[buffer_counter];
                i32 b1 = bp[buffer_counter + 1] < 0 ? -bp[buffer_counter + 1] : bp[buffer_counter + 1];
                if (a > mute_limit || b1 > mute_limit) { i = buffer_counter / 2; break; }
                crc = (crc * 3 + bp[buffer_counter]) * 3 + bp[buffer_counter + 1];
            }
        }
    }

    if (i != sample_count) {
        i64 sc = (flags & MONO_FLAG) > 0 ? sample_count : 2 * sample_count;
        buffer_counter = bufferStartPos;
        if (sc > 0 && (buffer_counter < 0 || buffer_counter + sc > buffer_len)) THROW(wpc);
        while (sc-- > 0) buffer[buffer_counter++] = 0;
        wps->mute_error = 1;
        i = sample_count;
    }

    fixup_samples(wpc, wps, buffer, buffer_len, i, bufferStartPos);

    if ((flags & FALSE_STEREO) > 0) {
        i64 dest_idx = i * 2, src_idx = i, cnt = i;
        if (cnt > 0) CHK(bufferStartPos, bufferStartPos + dest_idx);
        while (cnt-- > 0) {
            src_idx--;
            buffer[--dest_idx + bufferStartPos] = buffer[src_idx + bufferStartPos];
            buffer[--dest_idx + bufferStartPos] = buffer[src_idx + bufferStartPos];
        }
    }

    wps->sample_index += i;
    wps->crc = crc;
    return i;
}

static int check_crc_error(rd_context *wpc) /* UnpackUtils.cs:1414-1421 */
{
    WavpackStream *wps = &wpc->stream;
    return wps->crc != wps->wphdr.crc ||
           ((wps->wphdr.flags & FLOAT_DATA) == 0 && !wps->wvxbits.is_null && wps->crc_x != wps->crc_mvx);
}

/* ------------------------------------------------------------------ */
/* DsdUtils decode                                                     */
/* ------------------------------------------------------------------ */
static i64 decode_fast(rd_context *wpc, i32 *output, long buffer_len, i64 sample_count, i64 bufferStartPos) /* DsdUtils.cs:244-304 */
{
    WavpackStream *wps = &wpc->stream;
    dsds *d = &wps->dsd;
    i64 total_samples = sample_count;
    if ((wps->wphdr.flags & MONO_DATA) == 0) total_samples *= 2;

    while (total_samples-- > 0) {
        u32 mult, index, i;
        int code;
        int p0_index = d->p0 * MAX_DSD_BITS_VALUE;
        if (d->summed_probabilities[p0_index + 255] == 0) return 0;
        mult = (d->high - d->low) / d->summed_probabilities[p0_index + 255];
        if (mult == 0) {
            if (d->data->len - d->byteptr >= 4)
                for (i = 4; i > 0; i--) d->value = (d->value << 8) | ba_get(wpc, d->data, d->byteptr++);
            d->low = 0;
            d->high = 0xFFFFFFFFu;
            mult = d->high / d->summed_probabilities[p0_index + 255];
            if (mult == 0) return 0;
        }
        index = (d->value - d->low) / mult;
        if (index >= d->summed_probabilities[p0_index + 255]) return 0;
        {
            i64 li = (i64)d->value_lookup[d->p0] + index;
            if (li < 0 || li >= d->lookup_len) THROW(wpc);
            code = d->lookup_buffer[li];
        }
        if (bufferStartPos < 0 || bufferStartPos >= buffer_len) THROW(wpc);
        output[bufferStartPos++] = code;
        if (code > 0) d->low += d->summed_probabilities[p0_index + code - 1] * mult;
        d->high = d->low + d->probabilities[p0_index + code] * mult - 1;
        wps->crc += (i32)((u32)wps->crc << 1) + code;
        if ((wps->wphdr.flags & MONO_DATA) > 0)
            d->p0 = code & (d->history_bins - 1);
        else {
            d->p0 = d->p1;
            d->p1 = code & (d->history_bins - 1);
        }
        while (((d->high ^ d->low) & 0xFF000000u) == 0 && d->byteptr < d->data->len) {
            d->value = (d->value << 8) | ba_get(wpc, d->data, d->byteptr++);
            d->high = (d->high << 8) | 0xFF;
            d->low <<= 8;
        }
    }
    return sample_count;
}

static inline void dsd_high_bit(rd_context *wpc, dsds *d, DSDfilters *sp) /* one channel-bit, DsdUtils.cs:408-441 */
{
    int pp = (sp->value >> (DSD_PRECISION - DSD_PRECISION_USE)) & PTABLE_MASK;
    u32 split = d->low + ((d->high - d->low) >> 8) * ((u32)d->ptable[pp] >> 16);
    if (d->value <= split) {
        d->high = split;
        d->ptable[pp] += (DSD_UP - d->ptable[pp]) >> DSD_DECAY;
        sp->filter0 = -1;
    } else {
        d->low = split + 1;
        d->ptable[pp] += (DSD_DOWN - d->ptable[pp]) >> DSD_DECAY;
        sp->filter0 = 0;
    }
    while (((d->high ^ d->low) & 0xFF000000u) == 0 && d->byteptr < d->data->len) {
        d->value = (d->value << 8) | ba_get(wpc, d->data, d->byteptr++);
        d->high = (d->high << 8) | 0xFF;
        d->low <<= 8;
    }
    sp->value += sp->filter6 * 8;
    sp->bytei = (i32)((u32)sp->bytei << 1) | (sp->filter0 & 1);
    sp->factor += (((sp->value ^ sp->filter0) >> 31) | 1) & ((sp->value ^ (sp->value - (sp->filter6 * 16))) >> 31);
    sp->filter1 += ((sp->filter0 & DSD_VALUE_ONE) - sp->filter1) >> 6;
    sp->filter2 += ((sp->filter0 & DSD_VALUE_ONE) - sp->filter2) >> 4;
    sp->filter3 += (sp->filter2 - sp->filter3) >> 4;
    sp->filter4 += (sp->filter3 - sp->filter4) >> 4;
    sp->value = (sp->filter4 - sp->filter5) >> 4;
    sp->filter5 += sp->value;
    sp->filter6 += (sp->value - sp->filter6) >> 3;
    sp->value = sp->filter1 - sp->filter5 + ((sp->filter6 * sp->factor) >> 2);
}

static i64 decode_high(rd_context *wpc, i32 *output, long buffer_len, i64 sample_count, i64 bufferStartPos) /* DsdUtils.cs:391-493 */
{
    WavpackStream *wps = &wpc->stream;
    dsds *d = &wps->dsd;
    i64 total_samples = sample_count;
    int stereo = (wps->wphdr.flags & MONO_DATA) > 0 ? 0 : 1;
    DSDfilters *sp = d->filters;

    while (total_samples-- > 0) {
        int bitcount = 8;
        sp[0].value = sp[0].filter1 - sp[0].filter5 + ((sp[0].filter6 * sp[0].factor) >> 2);
        if (stereo) sp[1].value = sp[1].filter1 - sp[1].filter5 + ((sp[1].filter6 * sp[1].factor) >> 2);
        while (bitcount-- > 0) {
            dsd_high_bit(wpc, d, &sp[0]);
            if (!stereo) continue;
            dsd_high_bit(wpc, d, &sp[1]);
        }
        if (bufferStartPos < 0 || bufferStartPos >= buffer_len) THROW(wpc);
        wps->crc += (i32)((u32)wps->crc << 1) + (output[bufferStartPos++] = sp[0].bytei & 0xFF);
        sp[0].factor -= (sp[0].factor + 512) >> 10;
        if (stereo) {
            if (bufferStartPos < 0 || bufferStartPos >= buffer_len) THROW(wpc);
            wps->crc += (i32)((u32)wps->crc << 1) + (output[bufferStartPos++] = sp[1].bytei & 0xFF);
            sp[1].factor -= (sp[1].factor + 512) >> 10;
        }
    }
    return sample_count;
}

static i64 unpack_dsd_samples(rd_context *wpc, i32 *buffer, long buffer_len, i64 sample_count, int bufferStartPos) /* DsdUtils.cs:56-136 */
{
    WavpackStream *wps = &wpc->stream;
    u32 flags = wps->wphdr.flags;

    if (wps->sample_index + sample_count > wps->wphdr.block_index + wps->wphdr.block_samples &&
        (wps->wphdr.block_index + wps->wphdr.block_samples - wps->sample_index) < sample_count)
        sample_count = wps->wphdr.block_index + wps->wphdr.block_samples - wps->sample_index;

    if (wps->wphdr.block_index > wps->sample_index || (i64)wps->wphdr.block_samples < sample_count)
        wps->mute_error = 1;

    if (!wps->mute_error) {
        if (wps->dsd.mode == 0) {
            i64 total_samples = sample_count * ((flags & MONO_DATA) > 0 ? 1 : 2);
            i64 bsp = bufferStartPos;
            if (!wps->dsd.data) THROW(wpc); /* no ID_DSD_BLOCK seen yet: NullReferenceException in the reference */
            if (wps->dsd.data->len - wps->dsd.byteptr < total_samples) total_samples = wps->dsd.data->len - wps->dsd.byteptr;
            while (total_samples-- > 0) {
                if (bsp < 0 || bsp >= buffer_len) THROW(wpc);
                wps->crc += (i32)((u32)wps->crc << 1) + (buffer[bsp++] = ba_get(wpc, wps->dsd.data, wps->dsd.byteptr++));
            }
        } else if (wps->dsd.mode == 1) {
            if (decode_fast(wpc, buffer, buffer_len, sample_count, bufferStartPos) == 0) wps->mute_error = 1;
        } else if (wps->dsd.mode == 3) {
            if (decode_high(wpc, buffer, buffer_len, sample_count, bufferStartPos) == 0) wps->mute_error = 1;
        } else
            wps->mute_error = 1;

        if (wps->sample_index + sample_count == wps->wphdr.block_index + wps->wphdr.block_samples && !wps->mute_error &&
            wps->crc != wps->wphdr.crc)
            wps->mute_error = 1;
    }

    if (wps->mute_error) {
        i64 samples_to_null;
        if (wpc->reduced_channels == 1 || wpc->config.num_channels == 1 || (flags & MONO_FLAG) > 0)
            samples_to_null = sample_count;
        else
            samples_to_null = sample_count * 2;
        if (samples_to_null > buffer_len) THROW(wpc);
        while (samples_to_null > 0) buffer[--samples_to_null] = 0x55; /* from index 0, ignores bufferStartPos (quirk C-11) */
        wps->sample_index += sample_count;
        return sample_count;
    }

    if ((flags & FALSE_STEREO) > 0) {
        i64 dest_idx = sample_count * 2, src_idx = sample_count, cnt = sample_count;
        if (cnt > 0) CHK(bufferStartPos, bufferStartPos + dest_idx);
        while (cnt-- > 0) {
            src_idx--;
            buffer[--dest_idx + bufferStartPos] = buffer[src_idx + bufferStartPos];
            buffer[--dest_idx + bufferStartPos] = buffer[src_idx + bufferStartPos];
        }
    }
    wps->sample_index += sample_count;
    return sample_count;
}

/* ------------------------------------------------------------------ */
/* WavPackUtils.cs                                                     */
/* ------------------------------------------------------------------ */
static void read_next_header(rd_context *wpc) /* WavPackUtils.cs:600-671 (forward scan, <= 1 MiB) */
{
    MemStream *infile = &wpc->infile;
    WavpackHeader *wphdr = &wpc->stream.wphdr;
    uint8_t buffer[32];
    i64 bytes_skipped = 0;
    int bleft = 0, counter = 0;

    for (;;) {
        for (int i = 0; i < bleft; i++) buffer[i] = buffer[32 - bleft + i];
        counter = 0;
        int cnt = 32 - bleft;
        if (ms_read(infile, buffer + bleft, cnt) != cnt) {
            wphdr->error = 1;
            return;
        }
        bleft = 32;
        if (buffer[0] == 'w' && buffer[1] == 'v' && buffer[2] == 'p' && buffer[3] == 'k' && (buffer[4] & 1) == 0 && buffer[6] < 16 &&
            buffer[7] == 0 && buffer[9] == 4 && buffer[8] >= (MIN_STREAM_VERS & 0xff) && buffer[8] <= (MAX_STREAM_VERS & 0xff)) {
            wphdr->ckSize = (u32)((buffer[7] << 24) | (buffer[6] << 16) | (buffer[5] << 8) | buffer[4]);
            wphdr->version = (int16_t)((buffer[9] << 8) | buffer[8]);
            wphdr->total_samples = (i64)(((u64)buffer[11] << 32) | ((u64)buffer[15] << 24) | ((u64)buffer[14] << 16) | ((u64)buffer[13] << 8) | buffer[12]);
            wphdr->block_index = (i64)(((u64)buffer[10] << 32) | ((u64)buffer[19] << 24) | ((u64)buffer[18] << 16) | ((u64)buffer[17] << 8) | buffer[16]);
            wphdr->block_samples = ((u32)buffer[23] << 24) | (buffer[22] << 16) | (buffer[21] << 8) | buffer[20];
            wphdr->flags = ((u32)buffer[27] << 24) | (buffer[26] << 16) | (buffer[25] << 8) | buffer[24];
            wphdr->crc = (i32)(((u32)buffer[31] << 24) | (buffer[30] << 16) | (buffer[29] << 8) | buffer[28]);
            wphdr->error = 0;
            wphdr->stream_position = infile->pos - bleft;
            if (wphdr->average_block_size == 0)
                wphdr->average_block_size = wphdr->ckSize;
            else
                wphdr->average_block_size = (wphdr->average_block_size + wphdr->ckSize) / 2;
            return;
        } else {
            counter++;
            bleft--;
        }
        while (bleft > 0 && buffer[counter] != 'w') {
            counter++;
            bleft--;
        }
        bytes_skipped += counter;
        if (bytes_skipped > 1048576L) {
            wphdr->error = 1;
            return;
        }
    }
}

static void ctx_init(rd_context *wpc, const uint8_t *file, size_t len)
{
    init_tables();
    memset(wpc, 0, sizeof(*wpc));
    wpc->read_buffer.p = (uint8_t *)calloc(BITSTREAM_BUFFER_SIZE, 1);
    wpc->read_buffer.len = BITSTREAM_BUFFER_SIZE;
    wpc->read_buffer.refs = -1;
    wpc->infile.data = file;
    wpc->infile.len = (i64)len;
    /* WavpackStream(): wvbits = new Bitstream() (non-null, end==0); wvcbits/wvxbits null */
    wpc->stream.wvbits.buf = ba_new(BITSTREAM_BUFFER_SIZE);
    wpc->stream.wvcbits.is_null = 1;
    wpc->stream.wvxbits.is_null = 1;
}

/* WavpackOpenFileInput (WavPackUtils.cs:36-120) after `new WavpackContext()`: reads from the stream's current position */
static void open_input(rd_context *wpc, uint32_t flags)
{
    WavpackStream *wps = &wpc->stream;

    wpc->total_samples = -1;
    while (wps->wphdr.block_samples == 0) {
        read_next_header(wpc);
        if (wps->wphdr.error) {
            wpc->error_message = "not compatible with this version of WavPack file!";
            return;
        }
        if (wps->wphdr.block_samples > 0 && wps->wphdr.total_samples != 0xFFFFFFFFLL)
            wpc->total_samples = wps->wphdr.total_samples;
        if (!unpack_init(wpc)) return;
    }

    wpc->config.flags = wpc->config.flags & ~0xffLL;
    wpc->config.flags = wpc->config.flags | (wps->wphdr.flags & 0xff);
    wpc->config.bytes_per_sample = (int)((wps->wphdr.flags & BYTES_STORED) + 1);
    wpc->config.float_norm_exp = wps->float_norm_exp;
    wpc->config.bits_per_sample = (int)((wpc->config.bytes_per_sample * 8) - ((wps->wphdr.flags & SHIFT_MASK) >> SHIFT_LSB));
    if ((wpc->config.flags & FLOAT_DATA) > 0) {
        wpc->config.bytes_per_sample = 3;
        wpc->config.bits_per_sample = 24;
    }
    if (wpc->config.sample_rate == 0) {
        if (wps->wphdr.block_samples == 0 || (wps->wphdr.flags & SRATE_MASK) == SRATE_MASK)
            wpc->config.sample_rate = 44100;
        else
            wpc->config.sample_rate = sample_rates[(int)((wps->wphdr.flags & SRATE_MASK) >> SRATE_LSB)];
    }
    if (wpc->config.num_channels == 0) {
        wpc->config.num_channels = (wps->wphdr.flags & MONO_FLAG) > 0 ? 1 : 2;
        wpc->config.channel_mask = 0x5 - wpc->config.num_channels;
    }
    if ((flags & RD_OPEN_2CH_MAX) > 0 && (wps->wphdr.flags & FINAL_BLOCK) == 0)
        wpc->reduced_channels = (wps->wphdr.flags & MONO_FLAG) != 0 ? 1 : 2;
    if ((flags & RD_OPEN_2CH_MAX) == 0 && wpc->config.num_channels > 2) {
        wpc->error_message = "only two channels supported!";
        return;
    }
    if ((wps->wphdr.flags & DSD_FLAG) != 0) {
        wpc->config.bytes_per_sample = 1;
        wpc->config.bits_per_sample = 8;
    }
}

rd_context *rd_open(const uint8_t *file, size_t len, uint32_t flags) /* WavPackUtils.cs:36-120 */
{
    rd_context *wpc = (rd_context *)malloc(sizeof(rd_context));
    ctx_init(wpc, file, len);
    if (setjmp(wpc->jb)) {
        wpc->error_message = "exception";
        return wpc;
    }
    open_input(wpc, flags);
    return wpc;
}

static void stream_release(WavpackStream *s) /* what the garbage collector does to a dropped WavpackStream */
{
    if (!s->wvbits.is_null) ba_unref(s->wvbits.buf);
    if (!s->wvcbits.is_null) ba_unref(s->wvcbits.buf);
    if (!s->wvxbits.is_null) ba_unref(s->wvxbits.buf);
    dsd_free(&s->dsd);
}

static void ctx_release_rest(rd_context *c) /* everything of a context except its stream */
{
    md_set_data(&c->md, NULL);
    free(c->read_buffer.p);
    free(c->file_extension);
    free(c->header);
    free(c->trailer);
    free(c);
}

void rd_close(rd_context *c)
{
    if (!c) return;
    stream_release(&c->stream);
    ctx_release_rest(c);
}

long rd_unpack_samples(rd_context *wpc, int32_t *buffer, long buffer_len, long samples_in) /* WavPackUtils.cs:200-282 */
{
    WavpackStream *wps = &wpc->stream;
    i64 samples = samples_in;
    i64 samples_unpacked = 0, samples_to_unpack;
    int num_channels = wpc->config.num_channels;
    i64 bcounter = 0;
    i64 buf_idx = 0;
    int bytes_returned = 0;

    if (setjmp(wpc->jb)) return -2;

    while (samples > 0) {
        if (wps->wphdr.block_samples == 0 || (wps->wphdr.flags & INITIAL_BLOCK) == 0 ||
            wps->sample_index >= wps->wphdr.block_index + wps->wphdr.block_samples) {
            read_next_header(wpc);
            if (wps->wphdr.error) break;
            if (wps->wphdr.block_samples == 0 || wps->sample_index == wps->wphdr.block_index)
                if (!unpack_init(wpc)) break;
        }
        if (wps->wphdr.block_samples == 0 || (wps->wphdr.flags & INITIAL_BLOCK) == 0 ||
            wps->sample_index >= wps->wphdr.block_index + wps->wphdr.block_samples)
            continue;

        if (wps->sample_index < wps->wphdr.block_index) {
            samples_to_unpack = wps->wphdr.block_index - wps->sample_index;
            if (samples_to_unpack > samples) samples_to_unpack = samples;
            wps->sample_index += samples_to_unpack;
            samples_unpacked += samples_to_unpack;
            samples -= samples_to_unpack;
            if (wpc->reduced_channels > 0) samples_to_unpack *= wpc->reduced_channels;
            else samples_to_unpack *= num_channels;
            bcounter = buf_idx;
            while (samples_to_unpack-- > 0) {
                if (bcounter < 0 || bcounter >= buffer_len) THROW(wpc);
                buffer[bcounter++] = 0;
            }
            buf_idx = bcounter;
            continue;
        }

        samples_to_unpack = wps->wphdr.block_index + wps->wphdr.block_samples - wps->sample_index;
        if (samples_to_unpack > samples) samples_to_unpack = samples;

        if ((wps->wphdr.flags & DSD_FLAG) > 0)
            unpack_dsd_samples(wpc, buffer, buffer_len, samples_to_unpack, (int)buf_idx);
        else
            unpack_samples(wpc, buffer, buffer_len, samples_to_unpack, (int)buf_idx);

        if (wpc->reduced_channels > 0)
            bytes_returned = (int)(samples_to_unpack * wpc->reduced_channels);
        else
            bytes_returned = (int)(samples_to_unpack * num_channels);
        buf_idx += bytes_returned;
        samples_unpacked += samples_to_unpack;
        samples -= samples_to_unpack;

        if (wps->sample_index == wps->wphdr.block_index + wps->wphdr.block_samples)
            if (check_crc_error(wpc)) wpc->crc_errors++;
        if (wps->sample_index == wpc->total_samples) break;
    }
    return (long)samples_unpacked;
}

/* seek (WavPackUtils.cs:521-594), reached through SetSample / SetTime (WavPackUtils.cs:502-512).
 * Returns 1 / 0 like the C# bool; -2 where the C# code would throw out of the call (anything but the IOException it
 * catches: divide by zero on a zero-length block, an IndexOutOfRange inside the nested open or unpack); -3 where it would
 * never return (`index -= toUnpack` with WavpackUnpackSamples stuck at 0, WavPackUtils.cs:573-578). */
static int seek_impl(rd_context *wpc, i64 targetSample)
{
    WavpackStream *wps = &wpc->stream;
    MemStream *infile = &wpc->infile;

    if (targetSample >= wpc->total_samples) return 0;
    if (targetSample < 0) targetSample = 0;

    int steps = 25;      /* maximum steps to position */
    const int min = 5;   /* min count of block for seek forward by just read header */

    while (steps-- > 0) {
        i64 seek_pos = wps->wphdr.stream_position;

        if (targetSample <= (i64)wps->wphdr.block_samples)
            seek_pos = 0;
        else if (targetSample < wps->wphdr.block_index || targetSample > wps->wphdr.block_index + (i64)wps->wphdr.block_samples) {
            i64 distance = targetSample - wps->wphdr.block_index;
            /* int * uint promotes to long in C# */
            distance += distance > 0 ? (-1 * (i64)wps->wphdr.block_samples + 1) : (-2 * (i64)wps->wphdr.block_samples + 1);
            if (wps->wphdr.block_samples == 0) return -2; /* DivideByZeroException is not an IOException */
            i64 blocks = distance / (i64)wps->wphdr.block_samples;
            if (blocks >= 0 && blocks <= min)
                seek_pos = -1;
            else
                seek_pos += blocks * wps->wphdr.average_block_size;
            if (seek_pos >= infile->len) seek_pos = -1;
        }

        if (seek_pos != -1) {
            if (seek_pos < 0) return 0; /* Stream.Seek before the beginning throws IOException: caught, `return false` */
            infile->pos = seek_pos;
        }

        read_next_header(wpc);
        if (wps->wphdr.error) continue;

        if (steps == 0 || (targetSample >= wps->wphdr.block_index && targetSample < (wps->wphdr.block_index + (i64)wps->wphdr.block_samples))) {
            i64 index = targetSample - wps->wphdr.block_index;
            infile->pos = wps->wphdr.stream_position;
            /* WavpackContext c = WavpackOpenFileInput(infile); wpc.stream = c.stream;
             * c shares the caller's stream object (position included) and owns a fresh read_buffer, config and error state,
             * all of which are dropped: only its WavpackStream survives */
            rd_context *c = (rd_context *)malloc(sizeof(rd_context));
            ctx_init(c, infile->data, (size_t)infile->len);
            c->infile.pos = infile->pos;
            if (setjmp(c->jb)) { /* an exception inside the nested open leaves seek() uncaught */
                infile->pos = c->infile.pos;
                rd_close(c);
                return -2;
            }
            open_input(c, 0);
            infile->pos = c->infile.pos;
            stream_release(&wpc->stream);
            wpc->stream = c->stream;
            ctx_release_rest(c);
            i32 temp_buf[RD_SAMPLE_BUFFER_SIZE];
            while (index > 0) {
                i64 toUnpack = index < RD_SAMPLE_BUFFER_SIZE / rd_get_reduced_channels(wpc) ? index : RD_SAMPLE_BUFFER_SIZE / rd_get_reduced_channels(wpc);
                toUnpack = rd_unpack_samples(wpc, temp_buf, RD_SAMPLE_BUFFER_SIZE, (long)toUnpack);
                if (toUnpack == -2) return -2;
                if (toUnpack == 0) return -3;
                index -= toUnpack;
            }
            return 1;
        }

        if (seek_pos == -1) {
            infile->pos = wps->wphdr.stream_position + (i64)wps->wphdr.ckSize;
            steps--; /* sic: "do not account forward seek by headers" decrements once more */
        }
    }
    return 0;
}

int rd_set_sample(rd_context *wpc, long sample) /* WavPackUtils.cs:509-512 */
{
    return seek_impl(wpc, (i64)sample);
}

int rd_set_time(rd_context *wpc, long milliseconds) /* WavPackUtils.cs:504-507 */
{
    return seek_impl(wpc, (i64)milliseconds / 1000 * wpc->config.sample_rate);
}

int rd_format_samples(const int32_t *src, long samcnt, int bps, uint8_t *pcm, long pcm_len, int offset, int dsd) /* WavPackUtils.cs:288-341 */
{
    i32 temp;
    long counter = offset, counter2 = 0;
    i64 len = (i64)samcnt * bps;
    if (pcm == NULL || pcm_len < len + offset) return 0;
    switch (bps) {
    case 1:
        if (dsd) while (samcnt-- > 0) pcm[counter++] = (uint8_t)src[counter2++];
        else while (samcnt-- > 0) pcm[counter++] = (uint8_t)(0x00FF & (src[counter2++] + 128));
        break;
    case 2:
        while (samcnt-- > 0) { temp = src[counter2++]; pcm[counter++] = (uint8_t)temp; pcm[counter++] = (uint8_t)(temp >> 8); }
        break;
    case 3:
        while (samcnt-- > 0) {
            temp = src[counter2++];
            pcm[counter++] = (uint8_t)temp; pcm[counter++] = (uint8_t)(temp >> 8); pcm[counter++] = (uint8_t)(temp >> 16);
        }
        break;
    case 4:
        while (samcnt-- > 0) {
            temp = src[counter2++];
            pcm[counter++] = (uint8_t)temp; pcm[counter++] = (uint8_t)(temp >> 8); pcm[counter++] = (uint8_t)(temp >> 16);
            /* SupportClass.URShift(temp, 24) (SupportClass.cs:26-32): low byte equals (temp >> 24) & 0xff */
            pcm[counter++] = (uint8_t)(temp >= 0 ? temp >> 24 : (temp >> 24) + (i32)((u32)2 << (~24 & 31)));
        }
        break;
    }
    return 1;
}

/* ---- getters, WavPackUtils.cs:133-499 ---- */
long rd_get_num_samples(rd_context *c, int native) { return (long)(native && c->dsd_multiplier > 0 ? c->total_samples * 8 : c->total_samples); }
long rd_get_sample_index(rd_context *c) { return (long)c->stream.sample_index; }
long rd_get_num_errors(rd_context *c) { return (long)c->crc_errors; }
int rd_lossy(rd_context *c) { return c->lossy_blocks || (c->config.flags & CONFIG_HYBRID_FLAG) != 0; }
long rd_get_sample_rate(rd_context *c)
{
    if (c->config.sample_rate != 0)
        return (long)(c->dsd_multiplier > 0 ? (i64)c->dsd_multiplier * c->config.sample_rate * 8 : c->config.sample_rate);
    return 44100;
}
int rd_get_num_channels(rd_context *c) { return c->config.num_channels != 0 ? c->config.num_channels : 2; }
int rd_get_bits_per_sample(rd_context *c)
{
    if (c->config.bits_per_sample != 0) return c->dsd_multiplier > 0 ? c->config.bits_per_sample / 8 : c->config.bits_per_sample;
    return 16;
}
int rd_get_bytes_per_sample(rd_context *c) { return c->config.bytes_per_sample != 0 ? c->config.bytes_per_sample : 2; }
int rd_get_reduced_channels(rd_context *c)
{
    if (c->reduced_channels != 0) return c->reduced_channels;
    if (c->config.num_channels != 0) return c->config.num_channels;
    return 2;
}
int rd_get_file_format(rd_context *c) { return c->file_format; }
const char *rd_get_file_extension(rd_context *c) { return c->file_extension ? c->file_extension : "wav"; }
const char *rd_get_error_message(rd_context *c) { return c->error_message; }
const uint8_t *rd_get_header(rd_context *c, long *len) { if (len) *len = c->header ? c->header_len : -1; return c->header; }
const uint8_t *rd_get_trailer(rd_context *c, long *len) { if (len) *len = c->trailer ? c->trailer_len : -1; return c->trailer; }
int rd_get_is_five(rd_context *c) { return c->five; }
int rd_get_version(rd_context *c) { return c->stream.wphdr.version; }
int rd_get_is_float(rd_context *c) { return (c->config.flags & CONFIG_FLOAT_DATA) > 0; }

int rd_get_mode(rd_context *wpc) /* WavPackUtils.cs:133-167 */
{
    int mode = 0;
    if ((wpc->config.flags & CONFIG_HYBRID_FLAG) != 0) mode |= MODE_HYBRID;
    else if ((wpc->config.flags & CONFIG_LOSSY_MODE) == 0) mode |= MODE_LOSSLESS;
    if (wpc->lossy_blocks) mode &= ~MODE_LOSSLESS;
    if ((wpc->config.flags & CONFIG_FLOAT_DATA) != 0) mode |= MODE_FLOAT;
    if ((wpc->config.flags & CONFIG_HIGH_FLAG) != 0) {
        mode |= MODE_HIGH;
        if ((wpc->config.flags & CONFIG_VERY_HIGH_FLAG) > 0 || (wpc->stream.wphdr.version < 0x405)) mode |= MODE_VERY_HIGH;
    }
    if ((wpc->config.flags & CONFIG_FAST_FLAG) != 0) mode |= MODE_FAST;
    if ((wpc->config.flags & CONFIG_EXTRA_MODE) != 0) mode |= MODE_EXTRA | ((wpc->config.xmode << 12) & MODE_XMODE);
    if (wpc->dsd_multiplier > 0) mode |= MODE_DSD;
    return mode;
}

void rd_get_compression_level(rd_context *wpc, char *out, size_t cap) /* WavPackUtils.cs:169-187 */
{
    int mode = rd_get_mode(wpc);
    const char *base = NULL;
    if ((mode & MODE_FAST) > 0) base = "Fast";
    else if ((mode & MODE_VERY_HIGH) > 0) base = "Very High";
    else if ((mode & MODE_HIGH) > 0) base = "High";
    if ((mode & MODE_EXTRA) > 0)
        snprintf(out, cap, "%s, Extra-%d", base ? base : "Default", (mode & MODE_XMODE) >> 12);
    else
        snprintf(out, cap, "%s", base ? base : "");
}

/* ---- oracle-only helpers ---- */
int32_t rd_dbg_block_crc(rd_context *c) { return c->stream.crc; }
int rd_dbg_mute_error(rd_context *c) { return c->stream.mute_error; }
uint32_t rd_dbg_block_flags(rd_context *c) { return c->stream.wphdr.flags; }
int rd_dbg_check_crc_error(rd_context *c) { return check_crc_error(c); }

long rd_dbg_unpack_current_block(rd_context *wpc, int32_t *buffer, long buffer_len, long chunk)
{
    WavpackStream *wps = &wpc->stream;
    if (setjmp(wpc->jb)) return -2;
    int out_ch = (wps->wphdr.flags & MONO_FLAG) ? 1 : 2;
    i64 done = 0, total = wps->wphdr.block_samples;
    while (done < total) {
        i64 n = total - done < chunk ? total - done : chunk;
        if ((wps->wphdr.flags & DSD_FLAG) > 0)
            unpack_dsd_samples(wpc, buffer + done * out_ch, buffer_len - done * out_ch, n, 0);
        else
            unpack_samples(wpc, buffer + done * out_ch, buffer_len - done * out_ch, n, 0);
        done += n;
    }
    return (long)done;
}

long rd_decode_file_pcm(const uint8_t *file, size_t len, uint32_t open_flags, long chunk_samples, uint8_t *pcm, size_t pcm_cap,
                        size_t *pcm_len, long *crc_errors) /* WvDemo.cs:110-135 loop */
{
    rd_context *c = rd_open(file, len, open_flags);
    long total = 0;
    size_t at = 0;
    if (rd_get_error_message(c)) { rd_close(c); return -1; }
    int nch = rd_get_reduced_channels(c);
    int byteps = rd_get_bytes_per_sample(c);
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)chunk_samples * nch);
    for (;;) {
        long n = rd_unpack_samples(c, tmp, chunk_samples * nch, chunk_samples);
        if (n < 0) { total = n; break; }
        total += n;
        if (n > 0) {
            size_t bytes = (size_t)n * nch * byteps;
            if (at + bytes > pcm_cap) { total = -3; break; }
            /* WvDemo leaves the dsd argument false (WvDemo.cs:125) */
            rd_format_samples(tmp, n * nch, byteps, pcm + at, (long)(pcm_cap - at), 0, 0);
            at += bytes;
        }
        if (n == 0) break;
    }
    if (pcm_len) *pcm_len = at;
    if (crc_errors) *crc_errors = rd_get_num_errors(c);
    free(tmp);
    rd_close(c);
    return total;
}

/* ---- WvDemo.Main restated (WvDemo.cs:15-174): the bytes it writes to <input>.<ext> and its exit code ----------
 * out receives the output file.  Exit codes as the demo returns them: 0 ok; 1 open error, exception while
 * writing (incl. the DivideByZeroException of `total_unpacked_samples % loop_samples` when the file has fewer
 * than 100 * SAMPLE_BUFFER_SIZE samples, WvDemo.cs:113,136 -- raised after the first chunk has been written),
 * wrong sample count, or CRC errors.  Returns the number of bytes written, or -1 if out is too small. */
#define RD_SAMPLE_BUFFER_SIZE 4096 /* Defines.cs:18 */
static void le32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

long rd_wvdemo(const uint8_t *file, size_t len, uint8_t *out, size_t cap, int *exit_code)
{
    size_t at = 0;
    int rc = 0;
    long total_unpacked = 0;
    rd_context *c = rd_open(file, len, 0); /* WvDemo.cs:28 */
    if (rd_get_error_message(c)) { /* WvDemo.cs:41-46 */
        rd_close(c);
        if (exit_code) *exit_code = 1;
        return 0;
    }
    const int nch = rd_get_reduced_channels(c);       /* WvDemo.cs:48-53 */
    const int bits = rd_get_bits_per_sample(c);
    const int byteps = rd_get_bytes_per_sample(c);
    const int block_align = byteps * nch;
    const long total_samples = rd_get_num_samples(c, 1);
    const long sample_rate = rd_get_sample_rate(c);

    long hlen = 0;
    const uint8_t *hdr = rd_get_header(c, &hlen); /* WvDemo.cs:74-77: stored header unless the file is float */
    if (hdr && !rd_get_is_float(c)) {
        if (at + (size_t)hlen > cap) { rd_close(c); return -1; }
        memcpy(out + at, hdr, (size_t)hlen);
        at += (size_t)hlen;
    } else { /* WvDemo.cs:78-105: RIFF(12) + "fmt "(8) + WaveHeader(16) + "data"(8) */
        if (at + 44 > cap) { rd_close(c); return -1; }
        uint8_t *h = out + at;
        const uint32_t data_bytes = (uint32_t)((int64_t)total_samples * block_align);
        memcpy(h, "RIFF", 4);
        le32(h + 4, (uint32_t)(data_bytes + 2 * 8 + 16) + 4u); /* RiffChunkHeader.cs:16 adds 4 for "WAVE" */
        memcpy(h + 8, "WAVE", 4);
        memcpy(h + 12, "fmt ", 4);
        le32(h + 16, 16);
        le16(h + 20, 1);                                           /* FormatTag */
        le16(h + 22, (uint32_t)nch & 0xffff);                      /* (ushort)num_channels */
        le32(h + 24, (uint32_t)sample_rate);
        le32(h + 28, (uint32_t)((int64_t)sample_rate * block_align)); /* BytesPerSecond */
        le16(h + 32, (uint32_t)block_align & 0xffff);
        le16(h + 34, (uint32_t)bits & 0xffff);
        memcpy(h + 36, "data", 4);
        le32(h + 40, data_bytes);
        at += 44;
    }

    const long samples_unpack = RD_SAMPLE_BUFFER_SIZE;                                  /* WvDemo.cs:111 */
    const long loop_samples = total_samples / 100 / samples_unpack * samples_unpack;   /* WvDemo.cs:113 */
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * (size_t)samples_unpack * (size_t)(nch > 0 ? nch : 1));
    int threw = 0;
    for (;;) { /* WvDemo.cs:118-141 */
        long n = rd_unpack_samples(c, tmp, samples_unpack * nch, samples_unpack);
        if (n < 0) { threw = 1; break; } /* an exception out of the decoder lands in the catch-all, WvDemo.cs:148 */
        total_unpacked += n;
        if (n > 0) {
            const size_t bytes = (size_t)n * (size_t)block_align;
            if (at + bytes > cap) { free(tmp); rd_close(c); return -1; }
            /* pcm_buffer has samples_unpack * block_align bytes; dsd defaults to false (WvDemo.cs:125) */
            if (!rd_format_samples(tmp, n * nch, byteps, out + at, (long)(samples_unpack * block_align), 0, 0)) break;
            at += bytes;
        }
        if (loop_samples == 0) { threw = 1; break; } /* DivideByZeroException, WvDemo.cs:136 */
        if (n == 0) break;
    }
    free(tmp);
    if (threw) { /* WvDemo.cs:148-153: the using block closes the file with what was written so far */
        rd_close(c);
        if (exit_code) *exit_code = 1;
        return (long)at;
    }
    long tlen = 0;
    const uint8_t *trl = rd_get_trailer(c, &tlen); /* WvDemo.cs:143-145 */
    if (trl) {
        if (at + (size_t)tlen > cap) { rd_close(c); return -1; }
        memcpy(out + at, trl, (size_t)tlen);
        at += (size_t)tlen;
    }
    const long num_samples = rd_get_num_samples(c, 0); /* WvDemo.cs:157-162 */
    if (num_samples != -1 && total_unpacked != num_samples) rc = 1;
    else if (rd_get_num_errors(c) > 0) rc = 1; /* WvDemo.cs:164-169 */
    rd_close(c);
    if (exit_code) *exit_code = rc;
    return (long)at;
}
