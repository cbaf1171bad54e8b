// wvb_pcm.cuh -- per-thread WavPack PCM block decoder (one block per thread, 32 blocks per warp in lock step).
//
// Device restatement of the reference's PCM hot path:
//   metadata contents   UnpackUtils.cs:156-382, WordsUtils.cs:75-187, FloatUtils.cs:15-30
//   bit reader          BitsUtils.cs:15-146           -> 64-bit window refilled 32 bits at a time
//   get_words           WordsUtils.cs:272-511         -> decode_word<>
//   decorr_*_pass*      UnpackUtils.cs:688-1240       -> per-sample streaming form, state in
//                                                        shared memory laid out [slot][thread] (bank == lane, conflict free)
//   unpack_samples      UnpackUtils.cs:510-686        -> joint stereo, mute check, CRC, FALSE_STEREO
//   fixup_samples       UnpackUtils.cs:1251-1404, FloatUtils.cs:32-56
//   WavpackFormatSamples WavPackUtils.cs:288-341      -> fused into the store
// The code is written against a tiny "platform" layer (WVB_DEV, ld_u32, SMEM accessor) so that
// tests/emul can compile the very same decode function for the host and run it without a GPU.
// That emulation is a test harness only; the shipped library contains the CUDA instantiation only.
#pragma once
#include <stdint.h>

#include "../../include/wvb.h"
#include "wv_tables.h"
#include "wvb_grid.h"

#ifdef __CUDACC__
#define WVB_DEV __device__ __forceinline__
#define WVB_DEV_NOINLINE __device__ __noinline__
#define WVB_TABLE __device__ const
static __device__ __forceinline__ uint32_t wvb_ld_u32(const uint8_t *p) { return __ldg((const uint32_t *)p); }
static __device__ __forceinline__ uint8_t wvb_ld_u8(const uint8_t *p) { return __ldg(p); }
static __device__ __forceinline__ int wvb_ffs(uint32_t x) { return __ffs((int)x); }
static __device__ __forceinline__ int wvb_clz(uint32_t x) { return __clz((int)x); }
#define WVB_SYNCWARP() __syncwarp()
#ifdef WVB_SYNC_LESS
#define WVB_SYNCWARP_MID() ((void)0)
#else
#define WVB_SYNCWARP_MID() __syncwarp()
#endif
static __device__ __forceinline__ uint32_t wvb_warp_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
// the rounding term of apply_weight, kept in constant memory so that it is an operand of the multiply-add instead of
// being rebuilt in registers for every pass
static __device__ __constant__ long long wvb_k512 = 512;
#define WVB_K512 wvb_k512
#else
#include <string.h>
#define WVB_DEV inline
#define WVB_DEV_NOINLINE inline
#define WVB_TABLE static const
static inline uint32_t wvb_ld_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint8_t wvb_ld_u8(const uint8_t *p) { return *p; }
static inline int wvb_ffs(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int wvb_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#define WVB_SYNCWARP() ((void)0)
#define WVB_SYNCWARP_MID() ((void)0)
static inline uint32_t wvb_warp_max(uint32_t v) { return v; }
#define WVB_K512 512LL
#endif

namespace wvb {

// (Measured and rejected, rounds 2a and 2b: moving the adds of constants to the integer FMA pipe.  tools/probe/pipe_probe.cu
// shows the ALU pipe (IADD3, LOP3, SHF, SEL, ISETP, LEA) and the integer FMA pipe (IMAD) each issuing every other cycle,
// independently (0.50 warp instructions per cycle per scheduler alone, 0.98 alternating), and this decoder has 263 ALU
// against 135 FMA-pipe instructions per frame.  But x + c as IMAD needs the multiplier 1 in a vector register -- the
// IMAD forms take one immediate, and not a uniform register together with one -- and at 80-90 registers the compiler
// re-loads that 1 from constant memory next to every use: 82.0 vs 78.3 ms.)
WVB_TABLE uint8_t k_log2[256] = WV_LOG2_TABLE_INIT;
WVB_TABLE uint8_t k_exp2[256] = WV_EXP2_TABLE_INIT;

enum : uint32_t {
    F_BYTES_STORED = 3, F_MONO = 4, F_HYBRID = 8, F_JOINT = 0x10, F_FLOAT = 0x80, F_INT32 = 0x100, F_HYB_BITRATE = 0x200,
    F_HYB_BALANCE = 0x400, F_FALSE_STEREO = 0x40000000u, F_DSD = 0x80000000u
};

// ---- format math ----------------------------------------------------------------------------
WVB_DEV int exp2s(int log) // WordsUtils.cs:633-646 (long shifts, count masked to 6 bits)
{
    const bool neg = log < 0;
    if (neg) log = -log;
    const uint64_t value = (uint64_t)(k_exp2[log & 0xff] | 0x100);
    const int sh = log >> 8;
    const int r = sh <= 9 ? (int)(value >> (9 - sh)) : (int)(value << ((sh - 9) & 63));
    return neg ? -r : r;
}

WVB_DEV int mylog2(uint32_t v32) // WordsUtils.cs:588-608, for 0 <= value < 2^32
{
    uint64_t v = (uint64_t)v32 + (v32 >> 9);
    if (v < 256) {
        const int dbits = 32 - wvb_clz((uint32_t)v);
        return (dbits << 8) + k_log2[(uint32_t)(v << (9 - dbits)) & 0xff];
    }
    const int dbits = v >> 32 ? 33 : 32 - wvb_clz((uint32_t)v);
    return (dbits << 8) + k_log2[(uint32_t)(v >> (dbits - 9)) & 0xff];
}

WVB_DEV int restore_weight(int w8) // WordsUtils.cs:653-661, w8 already sign-extended from sbyte
{
    int r = w8 << 3;
    if (r > 0) r += (r + 64) >> 7;
    return r;
}

// ---- bit reader ------------------------------------------------------------------------------
// LSB-first reader over [start, start+len); every bit past the end reads as 1 (the reference floods its
// buffer with 0xFF on overrun, BitsUtils.cs:132-146).  A plain LSB-first window is bit-identical to
// getbit/getbits.
struct BitReader {
    const uint8_t *base; // 4-byte aligned address of stream word 0
    uint32_t idx;        // next word to fetch
    uint32_t full_end;   // words [0, full_end) lie wholly inside the stream
    uint32_t tailw;      // word `full_end`: the stream's last 1..3 bytes with 0xFF above them (all ones if there are none)
    uint32_t w0, w1;     // window: bit `pos` of w1:w0 is the next bit of the stream; 64 - pos bits are valid
    uint32_t nw;         // word fetched one refill ahead: its load latency overlaps the decode of the bits before it
    int pos;             // 0..63 (every consumer leaves at least one valid bit or refills first)

    WVB_DEV uint32_t load_word()
    {
        uint32_t x = idx == full_end ? tailw : 0xFFFFFFFFu;
        if (idx < full_end) x = wvb_ld_u32(base + 4 * (size_t)idx);
        ++idx;
        return x;
    }
    WVB_DEV void init(const uint8_t *s, uint32_t len)
    {
        const uint32_t mis = (uint32_t)((uintptr_t)s & 3);
        base = s - mis;
        const uint32_t total = mis + len, r = total & 3u;
        full_end = total >> 2;
        tailw = r ? (wvb_ld_u32(base + 4 * (size_t)full_end) | (0xFFFFFFFFu << (8 * r))) : 0xFFFFFFFFu;
        idx = 0;
        w0 = load_word();
        w1 = load_word();
        nw = load_word();
        pos = 8 * (int)mis;
#ifdef __CUDA_ARCH__
        // Read the three words once before the sample loop.  All stream loads share one hardware scoreboard; a load still
        // formally pending on w0/w1 at loop entry makes the first peek of EVERY iteration wait on that scoreboard, i.e.
        // on the prefetch issued just before it.
        // (the ballot's result feeds a condition that never holds, only so that the read is not optimised away)
        uint32_t seen;
        // (active mask, not the full one: the WVX reader is set up under per-lane conditions)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tvote.sync.ballot.b32 %0, p, %2;\n\t}" : "=r"(seen) : "r"(w0 | w1 | nw), "r"(__activemask()));
        if (seen == 0x5a5a5a5au && len == 0xffffffffu) ++pos;
#endif
    }
    WVB_DEV void refill() // afterwards pos < 32: at least 33 valid bits
    {
        if (pos >= 32) {
#ifdef __CUDA_ARCH__
            // Written out so that the load lands in `nw` itself.  Left to the compiler, the fetched word goes to a
            // temporary that is copied into `nw` at once, which waits out the whole load latency at every refill (ncu:
            // 1.3-1.8 long-scoreboard stall cycles per issued instruction) instead of overlapping it with the next word.
            asm volatile("{\n\t"
                         ".reg .pred p, q;\n\t"
                         ".reg .u64 a;\n\t"
                         "mov.u32 %0, %1;\n\t"
                         "mov.u32 %1, %2;\n\t"
                         "setp.eq.u32 q, %3, %5;\n\t"
                         "selp.u32 %2, %6, 0xffffffff, q;\n\t"
                         "setp.lt.u32 p, %3, %5;\n\t"
                         "mad.wide.u32 a, %3, 4, %7;\n\t"
                         "@p ld.global.nc.u32 %2, [a];\n\t"
                         "add.u32 %3, %3, 1;\n\t"
                         "sub.s32 %4, %4, 32;\n\t"
                         "}"
                         : "+r"(w0), "+r"(w1), "+r"(nw), "+r"(idx), "+r"(pos)
                         : "r"(full_end), "r"(tailw), "l"(base));
#else
            w0 = w1;
            w1 = nw;
            nw = load_word();
            pos -= 32;
#endif
        }
    }
    WVB_DEV void consume(int n) { pos += n; }
    WVB_DEV uint32_t peek() const { return (uint32_t)((((uint64_t)w1 << 32) | w0) >> pos); } // the next min(32, 64 - pos) bits
    WVB_DEV uint64_t peek64() const { return (((uint64_t)w1 << 32) | w0) >> pos; }            // the next 64 - pos bits
    WVB_DEV uint32_t getbit() { refill(); const uint32_t b = peek() & 1u; consume(1); return b; }
    WVB_DEV uint32_t getbits(int n) // 0 <= n <= 32
    {
        refill();
        const uint32_t v = n >= 32 ? peek() : (peek() & ((1u << n) - 1u));
        consume(n);
        return v;
    }
};

// ---- entropy decoder state (words_data.cs / entropy_data.cs) ---------------------------------
template <bool HYB> struct Words;
template <> struct Words<false> {
    int m[2][3];
    int hold; // 0 none, 1 holding_one, 2 holding_zero
    uint32_t zeros_acc;
};
template <> struct Words<true> {
    int m[2][3];
    int hold;
    uint32_t zeros_acc;
    int slow[2], errlim[2];
    int64_t bacc[2], bdelta[2];
};

template <bool STEREO> WVB_DEV void update_error_limit(Words<true> &w, uint32_t flags) // WordsUtils.cs:195-261
{
    int b0 = (int)((w.bacc[0] += w.bdelta[0]) >> 16);
    if (!STEREO) {
        if (flags & F_HYB_BITRATE) {
            const int sl0 = (w.slow[0] + 128) >> 8;
            w.errlim[0] = sl0 - b0 > -0x100 ? exp2s(sl0 - b0 + 0x100) : 0;
        } else
            w.errlim[0] = exp2s(b0);
    } else {
        int b1 = (int)((w.bacc[1] += w.bdelta[1]) >> 16);
        if (flags & F_HYB_BITRATE) {
            const int sl0 = (w.slow[0] + 128) >> 8, sl1 = (w.slow[1] + 128) >> 8;
            if (flags & F_HYB_BALANCE) {
                const int balance = (sl1 - sl0 + b1 + 1) >> 1;
                if (balance > b0) { b1 = b0 * 2; b0 = 0; }
                else if (-balance > b0) { b0 = b0 * 2; b1 = 0; }
                else { b1 = b0 + balance; b0 = b0 - balance; }
            }
            w.errlim[0] = sl0 - b0 > -0x100 ? exp2s(sl0 - b0 + 0x100) : 0;
            w.errlim[1] = sl1 - b1 > -0x100 ? exp2s(sl1 - b1 + 0x100) : 0;
        } else {
            w.errlim[0] = exp2s(b0);
            w.errlim[1] = exp2s(b1);
        }
    }
}

// gamma-style count shared by the zero-run and the unary escape (WordsUtils.cs:321-335, 391-405).
// returns false on the 33-ones EOF.
WVB_DEV bool read_gamma(BitReader &br, uint32_t &value)
{
    int cbits = 0;
    while (cbits < 33 && br.getbit()) ++cbits;
    if (cbits == 33) return false;
    if (cbits < 2) { value = (uint32_t)cbits; return true; }
    uint32_t mask = 1, acc = 0;
    for (; --cbits > 0; mask <<= 1)
        if (br.getbit()) acc |= mask;
    value = acc | mask;
    return true;
}

// read_code for ranges >= 2^30, where the reference's int/long mixing matters (WordsUtils.cs:546-570, quirk C-3)
WVB_DEV uint32_t read_code_wide(BitReader &br, uint32_t low, uint32_t range)
{
    const int bitcount = 32 - wvb_clz(range);
    const int64_t extras = (int64_t)(int32_t)(1u << (bitcount & 31)) - (int64_t)range - 1;
    int64_t code = (int64_t)br.getbits(bitcount - 1);
    code &= (int64_t)(int32_t)((1u << ((bitcount - 1) & 31)) - 1u);
    if (code >= extras) {
        code = (code << 1) - extras;
        if (br.getbit()) ++code;
    }
    return (uint32_t)((int64_t)low + code);
}

// One word of get_words for channel CH.  Returns false when the reference would `break` (EOF / error).
// Written single-exit (no early returns, no goto): 32 lanes run this in lock step on different streams and must
// reconverge at the end of every word, which the compiler only guarantees for structured control flow.
template <bool HYB, bool STEREO, int CH> WVB_DEV bool decode_word(BitReader &br, Words<HYB> &w, uint32_t flags, int &out)
{
    int *med = w.m[CH];
    int st = 1; // 1: a code follows; 0: the word is a zero of a run; -1: the reference's `break` (EOF / bad code)
    out = 0;
    if (((w.m[0][0] | w.m[1][0]) & ~1) == 0 && w.hold == 0) { // zero-run regime, WordsUtils.cs:304-352
        if (w.zeros_acc > 0) {
            if (--w.zeros_acc > 0) st = 0;
        } else {
            uint32_t z = 0;
            if (!read_gamma(br, z)) st = -1;
            else {
                w.zeros_acc = z;
                if (z > 0) {
                    w.m[0][0] = w.m[0][1] = w.m[0][2] = 0;
                    w.m[1][0] = w.m[1][1] = w.m[1][2] = 0;
                    st = 0;
                }
            }
        }
        if constexpr (HYB) {
            if (st == 0) w.slow[CH] -= (w.slow[CH] + 128) >> 8;
        }
    }

    if (st > 0) {
        // The hold flags, the unary count and the three medians are computed without branches: the 32 lanes of a warp
        // rarely agree on `ones`, so every branch here would be walked on both sides anyway, plus its (re)convergence cost.
        br.refill();
        const bool held0 = w.hold == 2; // a zero left over from the previous word: no unary prefix (WordsUtils.cs:354-358)
        int t = wvb_ffs(~br.peek()) - 1; // WordsUtils.cs:361-428
        int nb = t + 1;
        if (!held0 && (unsigned)t >= 16u) { // escape: 16 ones, then a zero and a gamma-coded count, or a 17th one = end of stream
            br.consume(16);
            nb = 0;
            t = 0;
            if (br.getbit()) st = -1;
            else {
                uint32_t v = 0;
                if (!read_gamma(br, v)) st = -1;
                else t = (int)v + 16;
            }
        }
        br.consume(held0 ? 0 : nb);
        const int ones = held0 ? 0 : (t >> 1) + (w.hold == 1 ? 1 : 0);
        w.hold = held0 ? 0 : 2 - (t & 1);

        { // after a failed escape (st < 0) this still runs, on a harmless t = 0: the block's entropy state is dead by then
            if constexpr (HYB && CH == 0) update_error_limit<STEREO>(w, flags); // WordsUtils.cs:430-431

            uint32_t low, high; // WordsUtils.cs:433-475
            {
                const int m0 = med[0], m1 = med[1], m2 = med[2];
                const int g0 = (m0 >> 4) + 1, g1 = (m1 >> 4) + 1, g2 = (m2 >> 4) + 1;
                const bool p1 = ones >= 1, p2 = ones >= 2, p3 = ones >= 3;
                // ones == 0 decrements m0, more increments it; m1 moves only from ones >= 1 (down at 1, up above), m2 from 2
                med[0] = m0 + ((m0 + (p1 ? 128 : 126)) >> 7) * (p1 ? 5 : -2);
                med[1] = m1 + ((m1 + (p2 ? 64 : 62)) >> 6) * (p1 ? (p2 ? 5 : -2) : 0);
                med[2] = m2 + ((m2 + (p3 ? 32 : 30)) >> 5) * (p2 ? (p3 ? 5 : -2) : 0);
                low = (p1 ? (uint32_t)g0 : 0u) + (p2 ? (uint32_t)g1 : 0u) + (p3 ? (uint32_t)((ones - 2) * g2) : 0u);
                high = low + (uint32_t)(p2 ? g2 : p1 ? g1 : g0) - 1u;
            }

            uint32_t mid, sign;
            bool lossless_code = true;
            if constexpr (HYB) lossless_code = w.errlim[CH] == 0;
            const uint32_t range = high - low;
            const int bitcount = 32 - wvb_clz(range);
            if (lossless_code && bitcount < 31) {
                // read_code (WordsUtils.cs:546-570) and the sign bit (494-497) out of one 32-bit look-ahead: at most
                // bitcount-1 code bits, one extra bit and the sign, 31 bits in all
                if (br.pos + bitcount > 62) br.refill(); // rare below 17-bit codes: pos <= 31 + 16 here
                const uint32_t pk = br.peek();
                const uint32_t extras = (1u << bitcount) - range - 1u;
                int nbits = bitcount > 0 ? bitcount - 1 : 0;
                uint32_t code = pk & ((1u << nbits) - 1u);
                if (range != 0 && code >= extras) {
                    code = (code << 1) - extras + ((pk >> nbits) & 1u);
                    ++nbits;
                }
                mid = low + code;
                sign = (pk >> nbits) & 1u;
                br.consume(nbits + 1);
            } else {
                if (lossless_code)
                    mid = read_code_wide(br, low, range);
                else { // WordsUtils.cs:477-492: bisect the interval down to the error limit, one bit per halving
                    uint32_t lim = 0;
                    if constexpr (HYB) lim = (uint32_t)w.errlim[CH];
                    // at most 32 halvings and the sign bit: all 33 come out of one look at the window (>= 33 valid bits after a
                    // refill) instead of a refill check, a peek and a position update per bit
                    br.refill();
                    const uint64_t win = br.peek64();
                    int nb = 0;
                    mid = (uint32_t)(((uint64_t)high + low + 1) >> 1);
                    while (high - low > lim) {
                        if ((win >> nb) & 1u) low = mid;
                        else high = mid - 1;
                        ++nb;
                        mid = (uint32_t)(((uint64_t)high + low + 1) >> 1);
                    }
                    sign = (uint32_t)(win >> nb) & 1u;
                    br.consume(nb + 1);
                }
                if (lossless_code) sign = br.getbit();
            }
            out = sign ? (int)~mid : (int)mid; // WordsUtils.cs:494-497
            if constexpr (HYB) {
                if (flags & F_HYB_BITRATE) // WordsUtils.cs:501-502
                    w.slow[CH] = w.slow[CH] - ((w.slow[CH] + 128) >> 8) + mylog2(mid);
            }
        }
    }
    return st >= 0;
}

// ---- decorrelation ---------------------------------------------------------------------------
WVB_DEV int apply_weight(int w, int s) { return (int)(((int64_t)w * (int64_t)s + WVB_K512) >> 10); }
WVB_DEV int upd_weight(int w, int delta, int s, int in) // UnpackUtils.cs:707-713
{
    const int sg = (s ^ in) >> 31; // -1 when the signs differ: w -= delta, else w += delta
    const int d = delta ^ sg;
#ifdef __CUDA_ARCH__
    // spelled out so that the two non-zero tests fold into one predicate on a single add (the compiler otherwise
    // builds the increment through a chain of selects: 9 instructions per weight instead of 6)
    asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %1, 0;\n\tsetp.ne.and.s32 p, %2, 0, p;\n\t@p sub.s32 %0, %0, %3;\n\t@p add.s32 %0, %0, %4;\n\t}"
        : "+r"(w) : "r"(s), "r"(in), "r"(sg), "r"(d));
#else
    if (s != 0 && in != 0) w += d - sg;
#endif
    return w;
}
WVB_DEV int upd_weight_clip(int w, int delta, int s, int in) // UnpackUtils.cs:776-785
{
    // cross-channel weights never leave [-1024, 1024] (restore_weight's range, and every update clips), so clipping both
    // sides equals the reference's one-sided clip
    const int sg = (s ^ in) >> 31;
    int t = w + (delta ^ sg) - sg;
    t = t < -1024 ? -1024 : t > 1024 ? 1024 : t;
    if (s != 0 && in != 0) w = t;
    return w;
}

// pass descriptor word: term+5 [0..4], delta [5..7], ring mask [8..10], slot base [16..31]
WVB_DEV uint32_t pack_pass(int term, int delta, int mask, int base) { return (uint32_t)(term + 5) | ((uint32_t)delta << 5) | ((uint32_t)mask << 8) | ((uint32_t)base << 16); }
WVB_DEV int ring_mask(int term) { return term > 8 ? 1 : term < 0 ? 0 : term <= 1 ? 0 : term <= 2 ? 1 : term <= 4 ? 3 : 7; }

// (Measured and rejected in round 2: a software-pipelined pass loop -- descriptor, weights and history of pass p+1 loaded
// before the arithmetic of pass p, so that the shared-memory round trips leave the a/b dependency chain.  16-term lists at
// 8-10 resident warps: 171 vs 152 ms; 5-term lists: 119 vs 100 ms.  The extra instructions cost more than the latency.)
// SM(i): word i of this thread's private shared-memory column
template <bool STEREO, class SMEM> WVB_DEV void decorr_frame(SMEM &SM, int nterms, uint32_t t, int &a, int &b)
{
    for (int p = 0; p < nterms; ++p) {
        const uint32_t d = (uint32_t)SM(p);
        const int term = (int)(d & 31u) - 5, delta = (int)((d >> 5) & 7u), mask = (int)((d >> 8) & 7u), base = (int)(d >> 16);
        if (STEREO) {
            int wA = SM(base), wB = SM(base + 1);
            const int hb = base + 2;
            if (term > 8) { // 17 / 18: UnpackUtils.cs:700-768, 958-1008
                const int i0 = hb + (int)((t - 1) & 1u), i1 = hb + (int)(t & 1u);
                int h0 = SM(i0), h1 = SM(i1);
                int s = term == 17 ? 2 * h0 - h1 : (3 * h0 - h1) >> 1;
                const int oa = a + apply_weight(wA, s);
                wA = upd_weight(wA, delta, s, a);
                SM(i1) = oa;
                h0 = SM(i0 + 2); h1 = SM(i1 + 2);
                s = term == 17 ? 2 * h0 - h1 : (3 * h0 - h1) >> 1;
                const int ob = b + apply_weight(wB, s);
                wB = upd_weight(wB, delta, s, b);
                SM(i1 + 2) = ob;
                a = oa; b = ob;
            } else if (term > 0) { // 1..8: UnpackUtils.cs:885-938, 1119-1148
                const int ir = hb + (int)((t - (uint32_t)term) & (uint32_t)mask), iw = hb + (int)(t & (uint32_t)mask);
                int s = SM(ir);
                const int oa = a + apply_weight(wA, s);
                wA = upd_weight(wA, delta, s, a);
                SM(iw) = oa;
                s = SM(ir + mask + 1);
                const int ob = b + apply_weight(wB, s);
                wB = upd_weight(wB, delta, s, b);
                SM(iw + mask + 1) = ob;
                a = oa; b = ob;
            } else if (term == -1) { // UnpackUtils.cs:771-804, 1011-1042
                const int s = SM(hb);
                const int oa = a + apply_weight(wA, s);
                wA = upd_weight_clip(wA, delta, s, a);
                const int ob = b + apply_weight(wB, oa);
                wB = upd_weight_clip(wB, delta, oa, b);
                SM(hb) = ob;
                a = oa; b = ob;
            } else if (term == -2) { // UnpackUtils.cs:807-843, 1045-1080
                const int s = SM(hb);
                const int ob = b + apply_weight(wB, s);
                wB = upd_weight_clip(wB, delta, s, b);
                const int oa = a + apply_weight(wA, ob);
                wA = upd_weight_clip(wA, delta, ob, a);
                SM(hb) = oa;
                a = oa; b = ob;
            } else { // -3: UnpackUtils.cs:846-882, 1083-1116
                const int sA = SM(hb), sB = SM(hb + 1);
                const int oa = a + apply_weight(wA, sA);
                wA = upd_weight_clip(wA, delta, sA, a);
                const int ob = b + apply_weight(wB, sB);
                wB = upd_weight_clip(wB, delta, sB, b);
                SM(hb) = ob;     // samples_A[0] = B output
                SM(hb + 1) = oa; // samples_B[0] = A output
                a = oa; b = ob;
            }
            SM(base) = wA;
            SM(base + 1) = wB;
        } else { // decorr_mono_pass, UnpackUtils.cs:1156-1240 (negative terms already folded to term & 7 at init)
            int wA = SM(base);
            const int hb = base + 1;
            int s, iw;
            if (term > 8) {
                const int i0 = hb + (int)((t - 1) & 1u);
                iw = hb + (int)(t & 1u);
                const int h0 = SM(i0), h1 = SM(iw);
                s = term == 17 ? 2 * h0 - h1 : (3 * h0 - h1) >> 1;
            } else {
                const int ir = hb + (int)((t - (uint32_t)term) & (uint32_t)mask);
                iw = hb + (int)(t & (uint32_t)mask);
                s = SM(ir);
            }
            const int oa = a + apply_weight(wA, s);
            wA = upd_weight(wA, delta, s, a);
            SM(iw) = oa;
            SM(base) = wA;
            a = oa;
        }
    }
}

template <bool STEREO, class SMEM> WVB_DEV void truncate_weights(SMEM &SM, int nterms) // (short) casts at UnpackUtils.cs:942-943,1152-1153,1239
{
    for (int p = 0; p < nterms; ++p) {
        const int base = (int)((uint32_t)SM(p) >> 16);
        SM(base) = (int)(int16_t)SM(base);
        if (STEREO) SM(base + 1) = (int)(int16_t)SM(base + 1);
    }
}


// ---- decorrelator policies -------------------------------------------------------------------
// GenericDecorr: any term list, state in the thread's shared-memory column (decorr_frame above).
template <bool STEREO> struct GenericDecorr {
    static constexpr bool kFixed = false;
    template <class SMEM> WVB_DEV bool load(SMEM &, int) { return true; }
    template <class SMEM> WVB_DEV void frame(SMEM &SM, int nterms, uint32_t t, int &a, int &b) { decorr_frame<STEREO>(SM, nterms, t, a, b); }
    template <class SMEM> WVB_DEV void truncate(SMEM &SM, int nterms) { truncate_weights<STEREO>(SM, nterms); }
};

// FixedDecorr: a compile-time term list (decoder order).  Weights and history live in registers: after unrolling every
// index below is a constant, so the arrays are scalarised.  History h[j] = x[t-1-j] is a shift register.
template <bool STEREO, int... TERMS> struct FixedDecorr {
    static constexpr bool kFixed = true;
    static constexpr int N = (int)sizeof...(TERMS);
    int wA[N], wB[N];
    uint64_t dl; // the N 3-bit deltas packed (N <= 21)
    int hA[N][8], hB[N][8];

    static constexpr int term_at(int p)
    {
        constexpr int T[N] = {TERMS...};
        return T[p];
    }
    static constexpr int depth(int term) { return term > 8 ? 2 : term < 0 ? 1 : term; }

    // pull the state the generic metadata parser left in shared memory; false if this block's terms differ from TERMS
    template <class SMEM> WVB_DEV bool load(SMEM &SM, int nterms)
    {
        bool match = nterms == N;
        dl = 0;
#pragma unroll
        for (int p = 0; p < N; ++p) {
            constexpr int T[N] = {TERMS...};
            const int term = T[p];
            const uint32_t d = match ? (uint32_t)SM(p) : 0u;
            if ((int)(d & 31u) - 5 != term) match = false;
            const int mask = (int)((d >> 8) & 7u), base = (int)(d >> 16);
            dl |= (uint64_t)((d >> 5) & 7u) << (3 * p);
            wA[p] = match ? SM(base) : 0;
            wB[p] = (STEREO && match) ? SM(base + 1) : 0;
            const int hb = base + (STEREO ? 2 : 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hA[p][j] = 0;
                hB[p][j] = 0;
                if (j < depth(term) && match) {
                    if (term < 0) { // -1: prev B out; -2: prev A out; -3: slot 0 = prev B out, slot 1 = prev A out
                        hA[p][0] = SM(hb);
                        if (term == -3) hB[p][0] = SM(hb + 1);
                    } else {
                        hA[p][j] = SM(hb + ((-1 - j) & mask)); // x[-1-j] lives in ring slot (-1-j) & mask
                        if (STEREO) hB[p][j] = SM(hb + mask + 1 + ((-1 - j) & mask));
                    }
                }
            }
        }
        return match;
    }

    template <int TERM> WVB_DEV static int predict(const int *h)
    {
        if (TERM == 17) return 2 * h[0] - h[1];
        if (TERM == 18) return (3 * h[0] - h[1]) >> 1;
        return h[TERM - 1];
    }
    template <int TERM> WVB_DEV static void push(int *h, int v)
    {
        constexpr int n = TERM > 8 ? 2 : TERM;
#pragma unroll
        for (int j = n - 1; j > 0; --j) h[j] = h[j - 1];
        h[0] = v;
    }

    template <int P, class SMEM> WVB_DEV void pass(int &a, int &b)
    {
        constexpr int T[N] = {TERMS...};
        constexpr int term = T[P];
        const int delta = (int)((uint32_t)(dl >> (3 * P)) & 7u);
        if (term > 0) {
            int s = predict<term>(hA[P]);
            const int oa = a + apply_weight(wA[P], s);
            wA[P] = upd_weight(wA[P], delta, s, a);
            push<term>(hA[P], oa);
            a = oa;
            if (STEREO) {
                s = predict<term>(hB[P]);
                const int ob = b + apply_weight(wB[P], s);
                wB[P] = upd_weight(wB[P], delta, s, b);
                push<term>(hB[P], ob);
                b = ob;
            }
        } else if (term == -1) {
            const int s = hA[P][0];
            const int oa = a + apply_weight(wA[P], s);
            wA[P] = upd_weight_clip(wA[P], delta, s, a);
            const int ob = b + apply_weight(wB[P], oa);
            wB[P] = upd_weight_clip(wB[P], delta, oa, b);
            hA[P][0] = ob;
            a = oa; b = ob;
        } else if (term == -2) {
            const int s = hA[P][0];
            const int ob = b + apply_weight(wB[P], s);
            wB[P] = upd_weight_clip(wB[P], delta, s, b);
            const int oa = a + apply_weight(wA[P], ob);
            wA[P] = upd_weight_clip(wA[P], delta, ob, a);
            hA[P][0] = oa;
            a = oa; b = ob;
        } else {
            const int sA = hA[P][0], sB = hB[P][0];
            const int oa = a + apply_weight(wA[P], sA);
            wA[P] = upd_weight_clip(wA[P], delta, sA, a);
            const int ob = b + apply_weight(wB[P], sB);
            wB[P] = upd_weight_clip(wB[P], delta, sB, b);
            hA[P][0] = ob;
            hB[P][0] = oa;
            a = oa; b = ob;
        }
    }
    template <int P, class SMEM> WVB_DEV void passes(int &a, int &b)
    {
        if constexpr (P < N) {
            pass<P, SMEM>(a, b);
            passes<P + 1, SMEM>(a, b);
        }
    }
    template <class SMEM> WVB_DEV void frame(SMEM &, int, uint32_t, int &a, int &b) { passes<0, SMEM>(a, b); }
    template <class SMEM> WVB_DEV void truncate(SMEM &, int)
    {
#pragma unroll
        for (int p = 0; p < N; ++p) {
            wA[p] = (int)(int16_t)wA[p];
            if (STEREO) wB[p] = (int)(int16_t)wB[p];
        }
    }
};


// ---- fixup (UnpackUtils.cs:1251-1404, FloatUtils.cs:32-56) -----------------------------------
struct Fixup {
    int mode;  // 0 shift only, 1 float, 2 int32+wvx, 3 int32 redundancy only (no wvx)
    int shift; // final shift (already & 0x1f)
    int lossy;
    int minv, maxv, mins, maxs;
    int sent, zeros, ones, dups, max_width;
    int fshift; // float shift in [-32, 32]
};

WVB_DEV int shl32(int v, int n) { return (int)((uint32_t)v << (n & 31)); }

WVB_DEV void fixup_init(Fixup &f, uint32_t flags, const uint8_t *int32_info, const uint8_t *float_info, bool wvx_present, int max_width)
{
    f.mode = 0;
    f.lossy = (flags & F_HYBRID) != 0;
    int shift = (int)((flags >> 13) & 0x1f);
    f.sent = f.zeros = f.ones = f.dups = 0;
    f.max_width = max_width;
    f.fshift = 0;
    if (flags & F_FLOAT) {
        int s = (int)float_info[2] - (int)float_info[3] + (int)float_info[1];
        f.fshift = s > 32 ? 32 : s < -32 ? -32 : s;
        f.mode = 1;
    } else if (flags & F_INT32) {
        int sent = int32_info[0], zeros = int32_info[1], ones = int32_info[2], dups = int32_info[3];
        if (wvx_present) {
            f.mode = 2;
            f.sent = sent; f.zeros = zeros; f.ones = ones; f.dups = dups;
        } else if (sent == 0 && (zeros + ones + dups) != 0) {
            while (f.lossy && (flags & F_BYTES_STORED) == 3 && shift < 8) {
                if (zeros > 0) zeros--;
                else if (ones > 0) ones--;
                else if (dups > 0) dups--;
                else break;
                shift++;
            }
            f.mode = 3;
            f.zeros = zeros; f.ones = ones; f.dups = dups;
        } else
            shift += zeros + sent + ones + dups;
    }
    shift &= 0x1f;
    f.shift = shift;
    f.minv = f.maxv = f.mins = f.maxs = 0;
    if (f.lossy) {
        switch (flags & F_BYTES_STORED) {
        case 0: f.minv = -128 >> shift; f.maxv = 127 >> shift; break;
        case 1: f.minv = -32768 >> shift; f.maxv = 32767 >> shift; break;
        case 2: f.minv = -8388608 >> shift; f.maxv = 8388607 >> shift; break;
        default: f.minv = (int)(0x80000000u >> shift); f.maxv = 0x7FFFFFFF >> shift; break; // quirk C-7
        }
        f.mins = shl32(f.minv, shift);
        f.maxs = shl32(f.maxv, shift);
    }
}

WVB_DEV int fixup_redundancy(const Fixup &f, int v)
{
    if (f.zeros != 0) v = shl32(v, f.zeros);
    else if (f.ones != 0) v = shl32(v + 1, f.ones) - 1;
    else if (f.dups != 0) v = shl32(v + (v & 1), f.dups) - (v & 1);
    return v;
}

WVB_DEV int fixup_value(const Fixup &f, int v, BitReader &wvx, int &crc_x)
{
    if (f.mode == 1) {
        if (f.fshift > 0) v = shl32(v, f.fshift);
        else if (f.fshift < 0) v = v >> ((-f.fshift) & 31);
        return v > 8388607 ? 8388607 : v < -8388608 ? -8388608 : v;
    }
    if (f.mode == 2) {
        const uint32_t mask = (f.sent & 31) ? ((1u << (f.sent & 31)) - 1u) : 0u; // (1U << sent_bits) - 1 with a masked count
        if (f.sent > 0) {
            if (f.max_width > 0) {
                const int pv = v < 0 ? ~v : v;
                const int width = (32 - wvb_clz((uint32_t)pv)) + f.sent;
                int btr = f.sent;
                if (width <= f.max_width || (btr -= width - f.max_width) > 0) {
                    const uint32_t data = wvx.getbits(btr > 32 ? 32 : btr) & mask;
                    v = shl32((int)((uint32_t)shl32(v, btr) | data), f.sent - btr);
                } else
                    v = shl32(v, f.sent);
            } else {
                const uint32_t data = wvx.getbits(f.sent > 32 ? 32 : f.sent) & mask;
                v = (int)(((uint32_t)v << (f.sent & 31)) | data);
            }
        }
        v = fixup_redundancy(f, v);
        crc_x = crc_x * 9 + (v & 0xffff) * 3 + ((v >> 16) & 0xffff);
    } else if (f.mode == 3)
        v = fixup_redundancy(f, v);
    if (f.lossy) return v < f.minv ? f.mins : v > f.maxv ? f.maxs : shl32(v, f.shift);
    return shl32(v, f.shift);
}

// ---- output ----------------------------------------------------------------------------------
WVB_DEV void store_unit(uint8_t *q, int v, int unit, int add128)
{
    if (unit == 4) *(int *)q = v;
    else if (unit == 2) *(uint16_t *)q = (uint16_t)v;
    else if (unit == 3) { q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); }
    else q[0] = (uint8_t)(v + add128);
}

// Packs a block's contiguous output bytes into aligned 32-bit stores.  One store instruction of a warp touches 32
// different output streams (32 sectors) whatever its width, so 24-bit stereo written byte by byte costs six such
// instructions per frame and saturates the load/store path (ncu: L1TEX 71 % busy, decoder issue rate 27 %); packed, it is 1.5.
// The first and last words of a block may be shared with its neighbours in the output: those are written bytewise.
struct OutWriter {
    uint8_t *p;   // 4-byte aligned address of the word being assembled
    uint64_t acc; // bytes not stored yet, from bit 0
    int fill;     // valid bits in acc (< 32 between calls)
    int skip;     // leading bytes of the first word that belong to whatever precedes the block; 0 once that word is out

    WVB_DEV void init(uint8_t *q)
    {
        const int mis = (int)((uintptr_t)q & 3);
        p = q - mis;
        acc = 0;
        fill = 8 * mis;
        skip = mis;
    }
    WVB_DEV void push(uint32_t v, int bits) // bits = 8, 16, 24 or 32; v already reduced to that width
    {
        acc |= (uint64_t)v << fill;
        fill += bits;
        if (fill >= 32) {
            const uint32_t word = (uint32_t)acc;
            if (skip) {
                for (int k = skip; k < 4; ++k) p[k] = (uint8_t)(word >> (8 * k));
                skip = 0;
            } else
                *(uint32_t *)p = word;
            p += 4;
            acc >>= 32;
            fill -= 32;
        }
    }
    WVB_DEV void finish() // the bytes of a last, partial word
    {
        for (int k = skip; k < (fill >> 3); ++k) p[k] = (uint8_t)(acc >> (8 * k));
    }
};

// ---- staged output of 16-bit stereo PCM (one aligned 32-bit word per frame) --------------------
// A store instruction of this decoder writes 32 different output streams, 4 bytes each.  L2 does not hold on to such
// partially written 32-byte sectors: ncu (round 2, profiles/r02_l2_write_path.txt) shows 52 % of the store sectors missing
// in L2 (8 stores per sector would miss once, 12.5 %), every miss filling the sector from DRAM and every eviction writing
// it back -- DRAM traffic 1.47x the algorithmic bytes, all of the excess on the output side.  So the words are staged in
// shared memory: each lane appends its word to a 16-word ring in its own column (a shared store in place of the global
// one), and every 8 frames reads one complete, 32-byte aligned sector of its own stream back and writes it with a single
// 256-bit store -- a warp instruction then writes 32 whole sectors.  Frames of all lanes advance together, so counted in
// output words from each stream's sector-aligned base (W = phase + t) the same sector index completes for every lane at
// the same iteration, whatever the stream's phase: the flush is warp-uniform control flow.  Only words of this block are
// written: a first or last sector shared with the neighbouring block in the slab, or cut short by a fault, goes out
// word by word.  (A first version moved half sectors between lanes to build the stores, two lanes per stream and
// st.global.v4: correct traffic, but 5-7 instructions per frame against 2 for this one.)
constexpr int STAGE_RING_SLOTS = 16;
#ifdef __CUDA_ARCH__
template <class SMEM> struct Stage16 {
    static constexpr uint32_t ROW = (uint32_t)SMEM::kThreads * 4u; // bytes between consecutive slots of a column
    static constexpr uint32_t RING = (uint32_t)STAGE_RING_SLOTS * ROW;
    char *ring;    // slot 0 of the ring rows (CTA-wide; this thread's word of slot k is at ring + k*ROW + 4*tid)
    uint2 *meta;   // this thread's {phase = W of frame 0, limit = first W not to write}
    uint8_t *base; // sector-aligned global address of the stream's word W = 0
    uint32_t x;    // ring offset of the slot the next frame goes to, with this thread's 4*tid folded into the low bits

    __device__ __forceinline__ void init(SMEM &SM, uint8_t *op, uint32_t n)
    {
        ring = (char *)(SM.origin + SM.ring_slot0() * SMEM::kThreads);
        meta = (uint2 *)(SM.stage_meta() + threadIdx.x);
        const uint64_t a = (uint64_t)(uintptr_t)op;
        const uint32_t phase = (uint32_t)(a >> 2) & 7u;
        base = (uint8_t *)(uintptr_t)(a & ~31ull);
        x = phase * ROW + 4u * threadIdx.x;
        *meta = make_uint2(phase, phase + n);
    }
    __device__ __forceinline__ void push(uint32_t word)
    {
        *(uint32_t *)(ring + x) = word;
        x = (x + ROW) & (RING - 1u); // (4*tid < ROW: the low bits pass through)
    }
    // frames t, t+1, ... of this lane are not written (fault / short get_words): the flush stops at what was produced
    __device__ __forceinline__ void stop_at(uint32_t t) { meta->y = meta->x + t; }
    // write out sector q (words 8q .. 8q+7) of this lane's stream
    __device__ __forceinline__ void flush(uint32_t q)
    {
        const uint2 m = *meta;
        const uint32_t lo = 8u * q;
        const char *src = ring + (x & (ROW - 1u)) + (lo & 15u) * ROW;
        uint8_t *dst = base + 32ull * q;
        if (lo >= m.x && lo + 8u <= m.y) {
            const uint32_t w0 = *(const uint32_t *)src, w1 = *(const uint32_t *)(src + ROW), w2 = *(const uint32_t *)(src + 2 * ROW),
                           w3 = *(const uint32_t *)(src + 3 * ROW), w4 = *(const uint32_t *)(src + 4 * ROW), w5 = *(const uint32_t *)(src + 5 * ROW),
                           w6 = *(const uint32_t *)(src + 6 * ROW), w7 = *(const uint32_t *)(src + 7 * ROW);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4),
                         "r"(w5), "r"(w6), "r"(w7)
                         : "memory");
        } else if (lo + 8u > m.x && lo < m.y) { // a sector shared with the neighbouring block, or the block's last words
#pragma unroll 1
            for (uint32_t i = 0; i < 8; ++i)
                if (lo + i >= m.x && lo + i < m.y) *(uint32_t *)(dst + 4 * i) = *(const uint32_t *)(src + i * ROW);
        }
    }
};
#endif

// The reference decodes a block in caller-sized pieces (one unpack_samples call each).  piece_bounds() recovers the piece
// [ps, pe) that contains sample t from the descriptor's call grid (wvb_grid.h); it is evaluated only at piece events and on
// faults, so the grid costs one register (the next event) in the sample loop.
WVB_DEV void piece_bounds(const wvb_block_desc &D, uint32_t n, uint32_t t, uint32_t &ps, uint32_t &pe) { piece_of(D, n, t, ps, pe); }
// next sample index at which the weights are cast to short: 8 samples into a stereo piece of >= 16 samples, and the piece end
template <bool STEREO> WVB_DEV uint32_t next_piece_event(uint32_t t, uint32_t ps, uint32_t pe)
{
    return (STEREO && pe - ps >= 16 && t < ps + 8) ? ps + 8 : pe;
}

// ---- the per-thread block decoder ------------------------------------------------------------
// STEREO: two coded channels (neither MONO_FLAG nor FALSE_STEREO).  HYB: HYBRID_FLAG.  GENFIX: float / int32 / hybrid fixup.
// F16: the block satisfies block_is_fast16 (wvb_plan.h; the planner gives such blocks their own launches); only meaningful
// for STEREO && !GENFIX.
template <bool STEREO, bool HYB, bool GENFIX, class SMEM, class DEC = GenericDecorr<STEREO>, bool F16 = false>
WVB_DEV void decode_block_pcm(SMEM &SM, const uint8_t *in, const wvb_block_desc &D, uint8_t *out, int out_format, wvb_block_result *res,
                              bool valid = true)
{
    const uint8_t *blk = in + D.in_offset;
    const uint32_t flags = D.flags;
    // lanes without a block (tail of the grid) and MUTE_ALL blocks stay in the warp with n == 0: every lane must reach
    // the warp-wide operations of the sample loop
    const bool mute_all = (D.bflags & WVB_BF_MUTE_ALL) != 0;
    uint32_t n = (valid && !mute_all) ? D.block_samples : 0;
    const int unit = out_format == WVB_OUT_INT32 ? 4 : (int)D.out_bps;
    const int add128 = (out_format == WVB_OUT_PCM && unit == 1) ? 128 : 0;
    const uint32_t frame_bytes = (uint32_t)unit * D.out_stride;
    const int out_ch = D.out_channels;
    uint8_t *op = out + D.out_offset + (uint32_t)unit * D.out_ch_offset;
    uint32_t rflags = 0;

    if (valid && D.gap_before) { // zero fill before the block (WavPackUtils.cs:227-251)
        uint8_t *g = out + D.out_offset - (uint64_t)D.gap_before * frame_bytes;
        for (uint32_t i = 0; i < D.gap_before; ++i, g += frame_bytes)
            for (int c = 0; c < D.out_stride; ++c) store_unit(g + c * unit, 0, unit, add128);
    }
    if (valid && mute_all) {
        uint8_t *q = op;
        for (uint32_t i = 0; i < D.block_samples; ++i, q += frame_bytes)
            for (int c = 0; c < out_ch; ++c) store_unit(q + c * unit, 0, unit, add128);
        res->crc = -1; res->crc_x = -1; res->mute_from = 0;
        res->rflags = WVB_RF_MUTED | WVB_RF_CRC_ERROR | WVB_RF_INEXACT;
    }
    if (D.bflags & WVB_BF_STALE_STATE) rflags |= WVB_RF_INEXACT;

    // ---- metadata contents -> state ----
    // The table is the caller's word: a term count or a state size beyond what this launch provides decodes nothing and is
    // reported (wvb_index never emits such a block; a hand-made or damaged table must not reach shared memory out of bounds)
    int nterms = (int)D.sub_len[WVB_SUB_TERMS];
    bool state_ok = nterms <= 16 && (int)D.smem_words <= SM.cap();
    if (!state_ok) nterms = 0;
    {
        const uint8_t *tp = blk + D.sub_off[WVB_SUB_TERMS];
        int base = nterms;
        for (int d = 0; d < nterms; ++d) { // decoder order d <-> file byte nterms-1-d (UnpackUtils.cs:170-181)
            const uint8_t tb = wvb_ld_u8(tp + (nterms - 1 - d));
            int term = (int)(tb & 0x1f) - 5;
            const int delta = (tb >> 5) & 7;
            if (!STEREO && term < 0) term &= 7; // decorr_mono_pass default branch (UnpackUtils.cs:1207)
            const int mask = ring_mask(term);
            SM(d) = (int)pack_pass(term, delta, mask, base);
            const int words = STEREO ? 2 + 2 * (mask + 1) : 1 + (mask + 1);
            if (base + words > SM.cap()) { state_ok = false; nterms = d; break; } // smem_words understated the need
            for (int k = 0; k < words; ++k) SM(base + k) = 0;
            base += words;
        }
        // weights (UnpackUtils.cs:196-239): file order walks decoder index nterms-1 downwards
        const uint8_t *wp = blk + D.sub_off[WVB_SUB_WEIGHTS];
        int cnt = (int)D.sub_len[WVB_SUB_WEIGHTS];
        if (STEREO) cnt >>= 1;
        // the index pass checks the count against the term count it CARRIES (like the reference); this block's own list can
        // be shorter or absent (weights without terms after a normal block: stale state), so clamp to what was laid out here
        if (cnt > nterms) cnt = nterms;
        for (int j = 0; j < cnt; ++j) {
            const int d = nterms - 1 - j;
            const int b0 = (int)((uint32_t)SM(d) >> 16);
            if (STEREO) {
                SM(b0) = (int)(int16_t)restore_weight((int8_t)wvb_ld_u8(wp + 2 * j));
                SM(b0 + 1) = (int)(int16_t)restore_weight((int8_t)wvb_ld_u8(wp + 2 * j + 1));
            } else
                SM(b0) = (int)(int16_t)restore_weight((int8_t)wvb_ld_u8(wp + j));
        }
        // history (UnpackUtils.cs:250-360) incl. quirk C-1: every entry is parsed with the term of pass nterms-1
        const uint8_t *sp = blk + D.sub_off[WVB_SUB_SAMPLES];
        const int slen = (int)D.sub_len[WVB_SUB_SAMPLES];
        if (slen > 0 && nterms > 0) {
            const int qterm = (int)(wvb_ld_u8(tp) & 0x1f) - 5; // raw term of decorr_passes[nterms-1] = first file byte
            int counter = 0;
            if (D.version == 0x402 && (flags & F_HYBRID)) counter += STEREO ? 4 : 2;
            int d = nterms - 1;
            int sA[8], sB[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) sA[k] = sB[k] = 0;
            while (counter < slen && d >= 0) {
#define RD16S(at) exp2s((int)(int16_t)((uint32_t)wvb_ld_u8(sp + (at)) | ((uint32_t)wvb_ld_u8(sp + (at) + 1) << 8)))
                if (qterm > 8) {
                    sA[0] = RD16S(counter); sA[1] = RD16S(counter + 2); counter += 4;
                    if (STEREO) { sB[0] = RD16S(counter); sB[1] = RD16S(counter + 2); counter += 4; }
                } else if (qterm < 0) {
                    sA[0] = RD16S(counter); sB[0] = RD16S(counter + 2); counter += 4;
                } else {
#pragma unroll
                    for (int m = 0; m < 8; ++m)
                        if (m < qterm) {
                            sA[m] = RD16S(counter); counter += 2;
                            if (STEREO) { sB[m] = RD16S(counter); counter += 2; }
                        }
                }
#undef RD16S
                // place the reference's samples_A/B[] image into this pass's ring according to ITS term
                const uint32_t dd = (uint32_t)SM(d);
                const int term = (int)(dd & 31u) - 5, mask = (int)((dd >> 8) & 7u), base0 = (int)(dd >> 16);
                const int hb = base0 + (STEREO ? 2 : 1);
                if (term > 8) {
                    SM(hb + 1) = sA[0]; SM(hb + 0) = sA[1]; // x[-1] -> slot 1, x[-2] -> slot 0
                    if (STEREO) { SM(hb + 3) = sB[0]; SM(hb + 2) = sB[1]; }
                } else if (term > 0) {
#pragma unroll
                    for (int m = 0; m < 8; ++m)
                        if (m < term) { // samples[m] = x[-term+m]
                            SM(hb + ((m - term) & mask)) = sA[m];
                            if (STEREO) SM(hb + mask + 1 + ((m - term) & mask)) = sB[m];
                        }
                } else if (term == -1) SM(hb) = sA[0];
                else if (term == -2) SM(hb) = sB[0];
                else { SM(hb) = sA[0]; SM(hb + 1) = sB[0]; }
                --d;
            }
        }
    }

    DEC dec;
    const bool terms_ok = dec.load(SM, nterms) && state_ok; // fixed-list kernels refuse (loudly) a block whose terms differ
    if (!terms_ok) n = 0;
    Words<HYB> w;
    {
        const uint8_t *ep = blk + D.sub_off[WVB_SUB_ENTROPY];
#define RD16U(p, at) ((int)((uint32_t)wvb_ld_u8((p) + (at)) | ((uint32_t)wvb_ld_u8((p) + (at) + 1) << 8)))
        w.m[0][0] = exp2s(RD16U(ep, 0)); w.m[0][1] = exp2s(RD16U(ep, 2)); w.m[0][2] = exp2s(RD16U(ep, 4));
        if (STEREO) { w.m[1][0] = exp2s(RD16U(ep, 6)); w.m[1][1] = exp2s(RD16U(ep, 8)); w.m[1][2] = exp2s(RD16U(ep, 10)); }
        else w.m[1][0] = w.m[1][1] = w.m[1][2] = 0;
        w.hold = 0;
        w.zeros_acc = 0;
        if constexpr (HYB) { // read_hybrid_profile, WordsUtils.cs:124-187
            Words<true> &h = w;
            h.slow[0] = h.slow[1] = 0; h.errlim[0] = h.errlim[1] = 0;
            h.bacc[0] = h.bacc[1] = 0; h.bdelta[0] = h.bdelta[1] = 0;
            const uint8_t *hp = blk + D.sub_off[WVB_SUB_HYBRID];
            const int hl = D.sub_off[WVB_SUB_HYBRID] ? (int)D.sub_len[WVB_SUB_HYBRID] : 0;
            if (D.sub_off[WVB_SUB_HYBRID]) {
                int k = 0;
                if (flags & F_HYB_BITRATE) {
                    h.slow[0] = exp2s(RD16U(hp, k)); k += 2;
                    if (STEREO) { h.slow[1] = exp2s(RD16U(hp, k)); k += 2; }
                }
                h.bacc[0] = (int64_t)(int32_t)((uint32_t)RD16U(hp, k) << 16); k += 2;
                if (STEREO) { h.bacc[1] = (int64_t)(int32_t)((uint32_t)RD16U(hp, k) << 16); k += 2; }
                if (k < hl) {
                    h.bdelta[0] = exp2s((int)(int16_t)RD16U(hp, k)); k += 2;
                    if (STEREO) { h.bdelta[1] = exp2s((int)(int16_t)RD16U(hp, k)); k += 2; }
                }
            }
        }
#undef RD16U
    }

    BitReader br;
    br.init(blk + D.sub_off[WVB_SUB_WV], D.sub_len[WVB_SUB_WV]);

    Fixup fx;
    BitReader wvx;
    int crc_x = -1, crc_mvx = 0;
    bool wvx_here = false;
    if (GENFIX) {
        int max_width = 0;
        wvx_here = (D.bflags & WVB_BF_WVX_PRESENT) && D.sub_off[WVB_SUB_WVX];
        if (wvx_here) { // init_wvx_bitstream, UnpackUtils.cs:115-147
            const uint8_t *xp = blk + D.sub_off[WVB_SUB_WVX];
            crc_mvx = (int)((uint32_t)wvb_ld_u8(xp) | ((uint32_t)wvb_ld_u8(xp + 1) << 8) | ((uint32_t)wvb_ld_u8(xp + 2) << 16) |
                            ((uint32_t)wvb_ld_u8(xp + 3) << 24));
            wvx.init(xp + 4, D.sub_len[WVB_SUB_WVX] - 4);
            if (D.bflags & WVB_BF_WVX_NEW) {
                if (flags & F_FLOAT) { wvx.getbits(5); wvx.getbits(5); }
                else max_width = (int)(wvx.getbits(5) & 0x1f);
            }
        } else
            wvx.init(blk, 0);
        fixup_init(fx, flags, D.int32_info, D.float_info, (D.bflags & WVB_BF_WVX_PRESENT) != 0, max_width);
        if ((flags & F_INT32) && (D.bflags & WVB_BF_WVX_PRESENT) && (flags & F_FALSE_STEREO)) rflags |= WVB_RF_INEXACT; // quirk C-6
    } else {
        fx.shift = (int)((flags >> 13) & 0x1f);
    }

    int mute_limit = (int)((1LL << ((flags >> 18) & 0x1f)) + 2); // UnpackUtils.cs:517
    if (flags & F_HYBRID) mute_limit *= 2;
    const bool joint = STEREO && (flags & F_JOINT);
    constexpr bool fast16 = F16 && STEREO && !GENFIX;
#ifdef __CUDA_ARCH__
    constexpr bool staged = fast16 && SMEM::kStaged; // through shared memory, full-sector stores (Stage16)
#else
    constexpr bool staged = false;                   // the host emulation (tests/emul) stores directly
#endif
    // the block's output is one contiguous byte run unless it is a channel pair inside wider frames (multichannel files)
    const bool packed = !fast16 && D.out_stride == out_ch && D.out_ch_offset == 0;
    const uint32_t unit_mask = unit == 4 ? 0xffffffffu : ((1u << (8 * unit)) - 1u);
    OutWriter ow;
    ow.init(op);

    // call/chunk grid: weights are cast to short at the end of every pass call, i.e. after the first 8 samples of a stereo
    // piece of >= 16 samples and at the end of the piece (UnpackUtils.cs:604-605,942-943,1152-1153,1239)
    uint32_t next_ev;
    {
        uint32_t ps, pe;
        piece_bounds(D, n, 0, ps, pe);
        next_ev = next_piece_event<STEREO>(0, ps, pe);
    }

    int crc = -1, crc_keep = 0;
    bool fault = false, eof_fault = false, crc_frozen = false;
    uint32_t fault_t = 0;
#ifdef __CUDA_ARCH__
    Stage16<SMEM> stage;
    if constexpr (staged) stage.init(SM, op, n);
#endif
    // All 32 lanes of a warp iterate together (warp-max trip count) and re-join at the top of every sample:
    // a lane left behind by a divergent branch must not be allowed to run the rest of its block alone.
    const uint32_t nmax = wvb_warp_max(n);
    uint32_t live_n = n; // drops to 0 when the lane faults: it then idles through the remaining iterations
    for (uint32_t t = 0; t < nmax; ++t) {
        WVB_SYNCWARP();
        const bool act = t < live_n;
        if (act) {
            if (t == next_ev) {
                if (!STEREO && crc_frozen) { crc = crc_keep; crc_frozen = false; } // (mono: every event is a piece boundary)
                dec.truncate(SM, nterms);
                uint32_t ps, pe;
                piece_bounds(D, n, t, ps, pe);
                next_ev = next_piece_event<STEREO>(t, ps, pe);
            }
            int a = 0, b = 0;
            if (!eof_fault) {
                // the second word is decoded even when the first came back short: the state it disturbs is never used again
                bool got = decode_word<HYB, STEREO, 0>(br, w, flags, a);
                if (STEREO) got &= decode_word<HYB, STEREO, 1>(br, w, flags, b);
                if (!got) {           // get_words came back short (WordsUtils.cs:323,383,393): the reference still runs the
                    eof_fault = true; // passes and the CRC over the rest of the chunk, reading whatever the caller's buffer
                    a = b = 0;        // held.  We model those stale entries as zeros (exact for silence, and a CRC
                                      // mismatch either way otherwise).
#ifdef __CUDA_ARCH__
                    if constexpr (staged) stage.stop_at(t);
#endif
                }
            }
            dec.frame(SM, nterms, t, a, b);
            if (joint) { b -= (a >> 1); a += b; } // UnpackUtils.cs:615 (App. E-9)
            const int aa = a < 0 ? -a : a, ab = b < 0 ? -b : b;
            bool stop = aa > mute_limit || (STEREO && ab > mute_limit);
            if constexpr (!STEREO) {
                // Quirk C-5 (UnpackUtils.cs:572-575): the mono scan records the BUFFER index of the offending sample, which
                // includes the samples the same call decoded before this block.  When that sum happens to equal the piece
                // length, `i != sample_count` does not fire: nothing is muted, the scan (magnitude test and CRC) simply ended
                // early, and the piece goes out as decoded.  The CRC resumes with the next piece.
                if (stop) {
                    if (crc_frozen) stop = false;
                    else {
                        uint32_t ps, pe;
                        piece_bounds(D, n, t, ps, pe);
                        const uint32_t before = call_lookback(D, ps) * (uint32_t)D.out_stride;
                        if (before != 0 && before + (t - ps) == pe - ps) {
                            crc_frozen = true;
                            crc_keep = crc;
                            stop = false;
                        }
                    }
                }
            }
            if (!stop) {
                crc = crc * 3 + a;
                if (STEREO) crc = crc * 3 + b;
                if (!eof_fault) {
                    if constexpr (fast16) {
                        const uint32_t word = ((uint32_t)shl32(a, fx.shift) & 0xffffu) | ((uint32_t)shl32(b, fx.shift) << 16);
#ifdef __CUDA_ARCH__
                        if constexpr (staged) stage.push(word);
                        else
#endif
                            *(uint32_t *)op = word;
                    } else {
                        int va, vb = 0;
                        if (GENFIX) {
                            va = fixup_value(fx, a, wvx, crc_x);
                            if (STEREO) vb = fixup_value(fx, b, wvx, crc_x);
                        } else {
                            va = shl32(a, fx.shift);
                            if (STEREO) vb = shl32(b, fx.shift);
                        }
                        if (packed) {
                            ow.push((uint32_t)(va + add128) & unit_mask, 8 * unit);
                            if (out_ch == 2) ow.push((uint32_t)((STEREO ? vb : va) + add128) & unit_mask, 8 * unit);
                        } else { // a channel pair (or one channel) inside a wider frame
                            store_unit(op, va, unit, add128);
                            if (out_ch == 2) store_unit(op + unit, STEREO ? vb : va, unit, add128); // FALSE_STEREO duplicates (UnpackUtils.cs:668-680)
                        }
                    }
                    op += frame_bytes;
                } else { // the modelled remainder of the chunk ends with the piece
                    uint32_t ps, pe;
                    piece_bounds(D, n, t, ps, pe);
                    stop = t + 1 == pe;
                }
            }
            if (stop) {
                fault = true;
                fault_t = t;
                live_n = 0;
#ifdef __CUDA_ARCH__
                if constexpr (staged) { if (!eof_fault) stage.stop_at(t); }
#endif
            }
        }
#ifdef __CUDA_ARCH__
        if constexpr (staged) { if ((t & 7u) == 7u) stage.flush(t >> 3); } // (t is warp-uniform: every lane gets here)
#endif
    }
#ifdef __CUDA_ARCH__
    if constexpr (staged) { // what the loop's flushes have not reached: at most 14 words per stream, in two sectors
        stage.flush(nmax >> 3);
        stage.flush((nmax >> 3) + 1u);
    }
#endif

    if (packed) ow.finish();
    if (crc_frozen) crc = crc_keep;

    if (fault) { // mute from the start of the caller chunk that contains the fault (UnpackUtils.cs:649-664, App. E-10)
        rflags |= WVB_RF_MUTED;
        uint32_t piece_start, piece_end;
        piece_bounds(D, n, fault_t, piece_start, piece_end);
        uint8_t *q = out + D.out_offset + (uint32_t)unit * D.out_ch_offset + (uint64_t)piece_start * frame_bytes;
        // the first muted chunk still runs through fixup_samples with zeros; only INT32 "ones" (and WVX data bits) make that non-zero
        for (uint32_t i = piece_start; i < n; ++i, q += frame_bytes) {
            int z = 0;
            if (GENFIX && i < piece_end) {
                if (fx.mode == 2 && fx.sent > 0) rflags |= WVB_RF_INEXACT;
                if (fx.mode == 2 || fx.mode == 3) z = fixup_redundancy(fx, 0);
                if (fx.mode != 1) z = fx.lossy ? (z < fx.minv ? fx.mins : z > fx.maxv ? fx.maxs : shl32(z, fx.shift)) : shl32(z, fx.shift);
            }
            for (int c = 0; c < out_ch; ++c) store_unit(q + c * unit, z, unit, add128);
        }
        if (valid) res->mute_from = piece_start;
    } else if (valid && !mute_all) // (tail lanes alias the last block's result slot: they must not store into it)
        res->mute_from = n;

    if (!valid || mute_all) return;
    if (!terms_ok) { // planner routed the block to a kernel specialised for another term list (hash collision): never silent
        res->crc = -1; res->crc_x = -1; res->mute_from = 0; res->rflags = WVB_RF_BAD_BLOCK;
        return;
    }
    // check_crc_error, UnpackUtils.cs:1414-1421
    if (crc != D.crc) rflags |= WVB_RF_CRC_ERROR;
    if (eof_fault) rflags |= WVB_RF_INEXACT; // crc decision modelled, see the sample loop
    if (GENFIX && !(flags & F_FLOAT) && (D.bflags & WVB_BF_WVX_PRESENT)) {
        if (!wvx_here) rflags |= WVB_RF_INEXACT; // crc_mvx left over from an earlier block
        else if (crc_x != crc_mvx) rflags |= WVB_RF_CRC_ERROR | WVB_RF_CRCX_ERROR;
    }
    res->crc = crc;
    res->crc_x = crc_x;
    res->rflags = rflags;
}

} // namespace wvb
