// wvb_dsd_core.cuh -- DSD block decoders (DsdUtils.cs) as platform-neutral per-block functions.
//
//   mode 0  raw bytes                     DsdUtils.cs:73-82      one block per thread
//   mode 1  "fast": byte range coder with per-history-bin probability tables
//                                         DsdUtils.cs:149-304    one block per WARP: tables in shared memory,
//                                         the symbol lookup (lookup_buffer[value_lookup[p0]+index] == number of
//                                         cumulative entries <= index) is a 32-lane search
//   mode 3  "high": bit range coder + adaptive 6-filter predictor, 256-entry adaptive ptable
//                                         DsdUtils.cs:321-493    one block per thread, ptable in shared memory [256][thread]
// Like wvb_pcm.cuh this compiles for the device (wvb_dsd.cuh) and, for tests only, for the host (tests/emul).
#pragma once
#include "wvb_pcm.cuh"

namespace wvb {

enum { DSD_UP = 0x010000FE, DSD_DOWN = 0x00010000 };

// planner key stored in wvb_block_desc.smem_words for DSD blocks: mode | history_bits << 4 | rate_i << 8
WVB_DEV int dsd_key_mode(uint32_t k) { return (int)(k & 15u); }
WVB_DEV int dsd_key_hbits(uint32_t k) { return (int)((k >> 4) & 15u); }
WVB_DEV int dsd_key_rate(uint32_t k) { return (int)((k >> 8) & 255u); }

// sequential byte source over the ID_DSD_BLOCK payload with the reference's `byteptr < data.Length` guards.
// One thread per stream.  The stream is held as two aligned 32-bit words (the one with byte `pos` and the next) plus a
// bit offset, the second word fetched one word ahead: the range decoder consumes a byte every few steps, and a dependent
// byte load per renormalisation would put an L2 round trip on the critical path.  take()/skip() hand out up to four
// bytes at once without a loop, so that 32 lanes renormalising by different amounts stay converged.
#ifdef __CUDA_ARCH__
static __device__ __forceinline__ uint32_t wvb_fshr(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }         // sh 0..31
static __device__ __forceinline__ uint32_t wvb_fshl_clamp(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_lc(lo, hi, sh); } // sh 0..32
static __device__ __forceinline__ uint32_t wvb_bswap(uint32_t x) { return __byte_perm(x, 0, 0x0123); }
#else
static inline uint32_t wvb_fshr(uint32_t lo, uint32_t hi, uint32_t sh) { return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline uint32_t wvb_fshl_clamp(uint32_t lo, uint32_t hi, uint32_t sh) { return sh == 0 ? hi : sh >= 32 ? lo : (hi << sh) | (lo >> (32 - sh)); }
static inline uint32_t wvb_bswap(uint32_t x) { return __builtin_bswap32(x); }
#endif

struct ByteReader {
    uint32_t pos, len;      // next byte of the payload, payload length
    uint32_t cur, nxt;      // the aligned word holding byte `pos`, and the following word
    uint32_t ahead;         // the word after nxt, requested one refill early: no consumer ever waits for a load just issued
    uint32_t sh;            // bit offset of byte `pos` inside cur: 0, 8, 16 or 24
    const uint8_t *wnext;   // aligned address of the word after `ahead`
    WVB_DEV uint32_t fetch()
    {
        const uint32_t w = wvb_ld_u32(wnext); // the slab is padded: reading up to 12 bytes past the payload is safe
        wnext += 4;
        return w;
    }
    WVB_DEV void init(const uint8_t *s, uint32_t n, uint32_t at)
    {
        len = n; pos = at;
        const uint8_t *a = s + at;
        const uint32_t mis = (uint32_t)((uintptr_t)a & 3);
        wnext = a - mis;
        sh = 8 * mis;
        cur = fetch();
        nxt = fetch();
        ahead = fetch();
    }
    WVB_DEV bool more() const { return pos < len; }
    WVB_DEV uint32_t left() const { return len - pos; }
    // the next four bytes with the FIRST one in the top bits: what four rounds of `v = (v << 8) | byte` assemble
    WVB_DEV uint32_t peek4() const { return wvb_bswap(wvb_fshr(cur, nxt, sh)); }
    WVB_DEV void skip(uint32_t k) // 0 <= k <= 4 bytes
    {
        pos += k;
        sh += 8 * k;
#ifdef __CUDA_ARCH__
        // predicated, and written out so that the load lands in `ahead` itself (see BitReader::refill: left to the compiler
        // the word goes through a temporary, and the copy out of it waits for the load)
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     "setp.ge.u32 p, %3, 32;\n\t"
                     "@p mov.u32 %0, %1;\n\t"
                     "@p mov.u32 %1, %2;\n\t"
                     "@p ld.global.nc.u32 %2, [%4];\n\t"
                     "@p add.u64 %4, %4, 4;\n\t"
                     "@p sub.u32 %3, %3, 32;\n\t"
                     "}"
                     : "+r"(cur), "+r"(nxt), "+r"(ahead), "+r"(sh), "+l"(wnext));
#else
        if (sh >= 32) { cur = nxt; nxt = ahead; ahead = fetch(); sh -= 32; }
#endif
    }
    WVB_DEV uint32_t get()
    {
        const uint32_t b = peek4() >> 24;
        skip(1);
        return b;
    }
};

struct DsdOut { // where the decoded bytes go
    uint8_t *op;          // first output unit of the block (channel offset applied)
    uint32_t frame_bytes; // bytes per complete output sample
    int unit;             // 4 (int32) or 1 (PCM byte)
    int add;              // +128 for WVB_OUT_PCM (WavpackFormatSamples dsd:false, WvDemo.cs:125), 0 for WVB_OUT_DSD_RAW / int32
    int coded_ch;         // 1 or 2 values per byte-time in the stream
    int out_ch;           // channels written (FALSE_STEREO: 2)
    WVB_DEV void put(uint32_t j, int code) const // j-th coded value of the block
    {
        const uint32_t frame = coded_ch == 2 ? (j >> 1) : j;
        const int c = coded_ch == 2 ? (int)(j & 1u) : 0;
        uint8_t *q = op + (uint64_t)frame * frame_bytes + (uint32_t)(c * unit);
        store_unit(q, code, unit, add);
        if (coded_ch == 1 && out_ch == 2) store_unit(q + unit, code, unit, add);
    }
};

WVB_DEV void dsd_out_init(DsdOut &o, const wvb_block_desc &D, uint8_t *out, int out_format)
{
    o.unit = out_format == WVB_OUT_INT32 ? 4 : 1;
    o.add = out_format == WVB_OUT_PCM ? 128 : 0;
    o.frame_bytes = (uint32_t)o.unit * D.out_stride;
    o.op = out + D.out_offset + (uint32_t)o.unit * D.out_ch_offset;
    o.coded_ch = (D.flags & (F_MONO | F_FALSE_STEREO)) ? 1 : 2;
    o.out_ch = D.out_channels;
}

// result bookkeeping shared by the three modes.  fail_at: coded-value index where decoding stopped (or total).
WVB_DEV void dsd_finish(const wvb_block_desc &D, wvb_block_result *res, int crc, bool failed, uint32_t fail_sample, uint32_t extra_flags)
{
    uint32_t rf = extra_flags;
    const uint32_t n = D.block_samples;
    uint32_t mute_from = n, pe;
    if (failed) { // decode_fast/decode_high returned 0 (DsdUtils.cs:85-91): that piece and all later ones are 0x55
        piece_of(D, n, fail_sample < n ? fail_sample : (n ? n - 1 : 0), mute_from, pe);
        if (n == 0) mute_from = 0;
        rf |= WVB_RF_MUTED | WVB_RF_CRC_ERROR;
        // a failing first piece that starts mid-call leaves partially decoded bytes + stale caller data behind (quirk C-11)
        if (call_lookback(D, mute_from)) rf |= WVB_RF_INEXACT;
    } else if (crc != D.crc) { // DsdUtils.cs:99-101: only the last piece is muted
        mute_from = 0;
        if (n) piece_of(D, n, n - 1, mute_from, pe);
        rf |= WVB_RF_MUTED | WVB_RF_CRC_ERROR;
    }
    res->crc = crc;
    res->crc_x = -1;
    res->mute_from = mute_from;
    res->rflags = rf;
}

// ---- mode 0 -----------------------------------------------------------------------------------
// A DSD block the reference decodes with state left over from an earlier block (WVB_BF_MUTE_ALL): zeros, then the last
// piece is muted by the second pass exactly as after a CRC failure.  Flagged inexact.
WVB_DEV void dsd_stale_block(const wvb_block_desc &D, uint8_t *out, int out_format, wvb_block_result *res, int lane, int nlanes)
{
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
    for (uint32_t j = (uint32_t)lane; j < total; j += (uint32_t)nlanes) o.put(j, 0);
    if (lane == 0) {
        wvb_block_desc tmp = D;
        tmp.crc = 0; // crc below is -1: always a mismatch
        dsd_finish(tmp, res, -1, false, 0, WVB_RF_INEXACT);
    }
}

WVB_DEV void dsd_decode_raw(const uint8_t *in, const wvb_block_desc &D, uint8_t *out, int out_format, wvb_block_result *res)
{
    if (D.bflags & WVB_BF_MUTE_ALL) { dsd_stale_block(D, out, out_format, res, 0, 1); return; }
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
    const uint32_t len = D.sub_len[WVB_SUB_DSD];
    uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
    if (len - 2 < total) total = len - 2; // DsdUtils.cs:77-78
    int crc = -1;
    for (uint32_t j = 0; j < total; ++j) {
        const int b = (int)wvb_ld_u8(p + 2 + j);
        crc = crc * 3 + b;
        o.put(j, b);
    }
    dsd_finish(D, res, crc, false, 0, 0);
}

// ---- mode 3 ("high") --------------------------------------------------------------------------
struct DsdFilt { int value, f0, f1, f2, f3, f4, f5, f6, factor, bytei; };

struct RangeDec {
    uint32_t low, high, value;
    ByteReader br;
    // DsdUtils.cs:424-429: `while (((high ^ low) & 0xFF000000) == 0 && more) { shift one byte in }`.  After j rounds the
    // top byte of high ^ low is byte j (from the top) of the ORIGINAL xor, and after four rounds the xor is all ones, so the
    // loop runs min(leading zero bytes of high ^ low, bytes left) times: one count, three funnel shifts, no branch.
    WVB_DEV void normalize()
    {
        uint32_t k = (uint32_t)wvb_clz(high ^ low) >> 3; // 0..4
        const uint32_t left = br.left();
        if (k > left) k = left;
        const uint32_t s8 = 8 * k;
        value = wvb_fshl_clamp(br.peek4(), value, s8);
        high = wvb_fshl_clamp(0xFFFFFFFFu, high, s8);
        low = wvb_fshl_clamp(0u, low, s8);
        br.skip(k);
    }
};

// adaptive probability table access: PT(i) for a plain int column; wvb_dsd.cuh overloads these for its packed column
template <class SMEM> WVB_DEV int pt_load(SMEM &PT, int i) { return PT(i); }
template <class SMEM> WVB_DEV void pt_store(SMEM &PT, int i, int v) { PT(i) = v; }

template <class SMEM> WVB_DEV void dsd_high_bit(SMEM &PT, RangeDec &rc, DsdFilt &s) // DsdUtils.cs:408-441
{
    const int pp = (s.value >> 8) & 255;
    int pt = pt_load(PT, pp);
    const uint32_t split = rc.low + ((rc.high - rc.low) >> 8) * ((uint32_t)pt >> 16);
    const bool one = rc.value <= split; // (selects, not branches: the lanes of a warp decode different streams)
    rc.high = one ? split : rc.high;
    rc.low = one ? rc.low : split + 1;
    pt += ((one ? (int)DSD_UP : (int)DSD_DOWN) - pt) >> 8;
    s.f0 = one ? -1 : 0;
    pt_store(PT, pp, pt);
    rc.normalize();
    s.value += s.f6 * 8;
    s.bytei = (int)((uint32_t)s.bytei << 1) | (s.f0 & 1);
    s.factor += (((s.value ^ s.f0) >> 31) | 1) & ((s.value ^ (s.value - (s.f6 * 16))) >> 31);
    s.f1 += ((s.f0 & (1 << 20)) - s.f1) >> 6;
    s.f2 += ((s.f0 & (1 << 20)) - s.f2) >> 4;
    s.f3 += (s.f2 - s.f3) >> 4;
    s.f4 += (s.f3 - s.f4) >> 4;
    s.value = (s.f4 - s.f5) >> 4;
    s.f5 += s.value;
    s.f6 += (s.value - s.f6) >> 3;
    s.value = s.f1 - s.f5 + ((s.f6 * s.factor) >> 2);
}

// PT(i): this thread's adaptive ptable entry i; ptable0: the 256 initial values for this block's rate_i (host-built,
// init_ptable DsdUtils.cs:321-341 is a pure function of rate_i because rate_s must be 20, App. E-7)
template <class SMEM>
WVB_DEV void dsd_decode_high(SMEM &PT, const int *ptable0, const uint8_t *in, const wvb_block_desc &D, uint8_t *out, int out_format,
                             wvb_block_result *res, bool valid)
{
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
    const bool stereo = o.coded_ch == 2;
    const uint32_t n = valid ? D.block_samples : 0;
    for (int i = 0; i < 256; ++i) pt_store(PT, i, ptable0[i]);
    RangeDec rc;
    rc.br.init(p, D.sub_len[WVB_SUB_DSD], 4); // rate_shift, mode, rate_i, rate_s
    DsdFilt sp[2];
    for (int c = 0; c < 2; ++c) {
        sp[c].value = sp[c].f0 = sp[c].f6 = sp[c].bytei = 0;
        sp[c].f1 = sp[c].f2 = sp[c].f3 = sp[c].f4 = sp[c].f5 = sp[c].factor = 0;
    }
    for (int c = 0; c < (stereo ? 2 : 1); ++c) { // DsdUtils.cs:365-378
        sp[c].f1 = (int)rc.br.get() << 12;
        sp[c].f2 = (int)rc.br.get() << 12;
        sp[c].f3 = (int)rc.br.get() << 12;
        sp[c].f4 = (int)rc.br.get() << 12;
        sp[c].f5 = (int)rc.br.get() << 12;
        int f = (int)rc.br.get();
        f |= (int)rc.br.get() << 8;
        sp[c].factor = (int)((uint32_t)f << 16) >> 16;
    }
    rc.high = 0xFFFFFFFFu;
    rc.low = 0;
    rc.value = 0;
    for (int i = 0; i < 4; ++i) rc.value = (rc.value << 8) | rc.br.get();

    int crc = -1;
    const uint32_t nmax = wvb_warp_max(n);
    for (uint32_t t = 0; t < nmax; ++t) {
        WVB_SYNCWARP();
        if (t < n) {
            sp[0].value = sp[0].f1 - sp[0].f5 + ((sp[0].f6 * sp[0].factor) >> 2);
            if (stereo) sp[1].value = sp[1].f1 - sp[1].f5 + ((sp[1].f6 * sp[1].factor) >> 2);
            for (int b = 0; b < 8; ++b) {
                dsd_high_bit(PT, rc, sp[0]);
                if (stereo) dsd_high_bit(PT, rc, sp[1]);
            }
            int v = sp[0].bytei & 0xFF;
            crc = crc * 3 + v;
            o.put(stereo ? 2 * t : t, v);
            sp[0].factor -= (sp[0].factor + 512) >> 10;
            if (stereo) {
                v = sp[1].bytei & 0xFF;
                crc = crc * 3 + v;
                o.put(2 * t + 1, v);
                sp[1].factor -= (sp[1].factor + 512) >> 10;
            }
        }
    }
    if (valid) dsd_finish(D, res, crc, false, 0, 0);
}

// ---- mode 1 ("fast") --------------------------------------------------------------------------
// Table for one block: summed[bins*256] u16 cumulative probabilities (the per-symbol probability is the difference of
// neighbours, so the reference's separate probabilities[] and lookup_buffer[] are not stored: 512 B per history bin).
// Lanes cooperate on the device; on the host nlanes == 1.
struct DsdFastTables {
    uint16_t *summed;
};

// Build the tables (DsdUtils.cs:157-229).  Control flow is uniform across the lanes of a warp: every lane walks the same
// bytes; `lane`/`nlanes` only split the stores and the prefix sums.  Returns the payload offset after the tables, or 0 if
// the reference would reject the block (the host index pass has already applied the same rules).
WVB_DEV uint32_t dsd_fast_build(const DsdFastTables &T, const uint8_t *p, uint32_t len, int lane, int nlanes, int &bins_out)
{
    uint32_t at = 2;
    const int history_bits = (int)wvb_ld_u8(p + at++);
    const int bins = 1 << history_bits;
    bins_out = bins;
    const uint32_t tot = 256u * (uint32_t)bins;
    const int max_probability = (int)wvb_ld_u8(p + at++);
    if (max_probability < 0xFF) {
        uint32_t outp = 0;
        while (outp < tot && at < len) {
            const int code = (int)wvb_ld_u8(p + at++);
            if (code > max_probability) {
                int z = code - max_probability;
                while (outp < tot && z-- > 0) {
                    if ((int)(outp % (uint32_t)nlanes) == lane) T.summed[outp] = 0;
                    outp++;
                }
            } else if (code != 0) {
                if ((int)(outp % (uint32_t)nlanes) == lane) T.summed[outp] = (uint16_t)code;
                outp++;
            } else
                break;
        }
        if (outp < tot) return 0;
        if (at < len) at++; // terminator byte (already validated == 0 by the index pass)
    } else {
        for (uint32_t i = (uint32_t)lane; i < tot; i += (uint32_t)nlanes) T.summed[i] = wvb_ld_u8(p + at + i);
        at += tot;
    }
    return at;
}

// after a barrier: in-place cumulative sums (summed[] holds the raw probabilities on entry), one bin at a time; each lane
// owns 256/nlanes consecutive symbols
template <class SCAN> WVB_DEV void dsd_fast_sums(const DsdFastTables &T, int bins, int lane, int nlanes, SCAN scan_exclusive)
{
    const int per = 256 / nlanes;
    for (int b = 0; b < bins; ++b) {
        uint32_t local = 0;
        for (int k = 0; k < per; ++k) local += T.summed[b * 256 + lane * per + k];
        uint32_t run = scan_exclusive(local);
        for (int k = 0; k < per; ++k) {
            run += T.summed[b * 256 + lane * per + k];
            T.summed[b * 256 + lane * per + k] = (uint16_t)run;
        }
    }
}

// The symbol loop of decode_fast (DsdUtils.cs:251-301) for ONE block per thread.  32 lanes run it in lock step on 32
// different blocks (warp-max trip count, __syncwarp per symbol, single exit), like the PCM sample loop.
// SUMOF(row, p0)       -> row[255], the total of history bin p0 (the device keeps it on chip)
// FIND(row, p0, index, below, cur) -> the decoded symbol = number of entries of row[0..255] that are <= index; also returns
//                         below = row[symbol-1] (0 for symbol 0) and cur = row[symbol]
// EMIT(j, code)        -> deliver the j-th decoded value;  DONE(count) is called once with the number of values delivered
template <class SUMOF, class FIND, class EMIT, class DONE>
WVB_DEV void dsd_fast_decode(const DsdFastTables &T, int bins, const uint8_t *p, uint32_t len, uint32_t at, bool mono, uint32_t total,
                             SUMOF sumof, FIND find, EMIT emit, DONE done, int &crc_out, bool &failed, uint32_t &fail_at)
{
    RangeDec rc;
    rc.br.init(p, len, at);
    rc.value = 0;
    for (int i = 0; i < 4; ++i) rc.value = (rc.value << 8) | rc.br.get(); // DsdUtils.cs:234-235 (>= 4 bytes validated on the host)
    rc.low = 0;
    rc.high = 0xFFFFFFFFu;
    int p0 = 0, p1 = 0, crc = -1;
    failed = false;
    fail_at = total;
    uint32_t delivered = 0;
    const uint32_t nmax = wvb_warp_max(total);
    for (uint32_t j = 0; j < nmax; ++j) {
        WVB_SYNCWARP();
        if (j < total && !failed) {
            const uint16_t *row = T.summed + p0 * 256;
            const uint32_t sum = sumof(row, p0);
            bool ok = sum != 0;
            uint32_t mult = 0, index = 0;
            if (ok) {
                mult = (rc.high - rc.low) / sum;
                if (mult == 0) {
                    if (rc.br.left() >= 4)
                        for (int i = 0; i < 4; ++i) rc.value = (rc.value << 8) | rc.br.get();
                    rc.low = 0;
                    rc.high = 0xFFFFFFFFu;
                    mult = rc.high / sum;
                    ok = mult != 0;
                }
            }
            if (ok) {
                index = (rc.value - rc.low) / mult;
                ok = index < sum;
            }
            if (ok) {
                uint32_t below, cur;
                const int code = find(row, p0, index, below, cur);
                emit(j, code);
                delivered = j + 1;
                rc.low += below * mult;
                rc.high = rc.low + (cur - below) * mult - 1;
                crc = crc * 3 + code;
                if (mono) p0 = code & (bins - 1);
                else { p0 = p1; p1 = code & (bins - 1); }
                rc.normalize();
            } else {
                failed = true;
                fail_at = j;
            }
        }
    }
    done(delivered);
    crc_out = crc;
}

// host-side table of initial ptables (mode 3), DsdUtils.cs:321-341
inline void dsd_init_ptable_host(int *table, int rate_i, int rate_s)
{
    int value = 0x808000, rate = rate_i << 8, c, i;
    for (c = (rate + 128) >> 8; c > 0; c--) value += (DSD_DOWN - value) >> 8;
    for (i = 0; i < 128; ++i) {
        table[i] = value;
        table[255 - i] = 0x100ffff - value;
        if (value > 0x010000) {
            rate += (rate * rate_s + 128) >> 8;
            for (c = (rate + 64) >> 7; c > 0; c--) value += (DSD_DOWN - value) >> 8;
        }
    }
}

} // namespace wvb
