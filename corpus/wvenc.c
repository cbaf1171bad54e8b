/*
 * corpus/wvenc.c -- synthetic WavPack corpus generator (see wvenc.h).
 *
 * Written from the stream format as the reference decoder parses it (SURVEY.md
 * Appendix A/B/D): every coding step below is the inverse of a decode step, with
 * the decode step's reference location cited.  Bench/test infrastructure only.
 */
#include "wvenc.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef int32_t i32;
typedef uint32_t u32;
typedef int64_t i64;
typedef uint64_t u64;

/* flag bits, SURVEY.md Appendix A (Defines.cs:28-44,86-94) */
#define F_MONO 4u
#define F_HYBRID 8u
#define F_JOINT 0x10u
#define F_FLOAT 0x80u
#define F_INT32 0x100u
#define F_HYB_BITRATE 0x200u
#define F_HYB_BALANCE 0x400u
#define F_INITIAL 0x800u
#define F_FINAL 0x1000u
#define F_FALSE_STEREO 0x40000000u
#define F_DSD 0x80000000u

/* ------------------------------------------------------------------ */
/* format math                                                          */
/* ------------------------------------------------------------------ */
#include "../wavpackdecoder_b200/csrc/wv_tables.h" /* generated closed-form log2/exp2 tables */
static const uint8_t T_LOG2[256] = WV_LOG2_TABLE_INIT, T_EXP2[256] = WV_EXP2_TABLE_INIT;
static int bitlen32(u32 v) { return v ? 32 - __builtin_clz(v) : 0; }
static void init_tables(void) {}
int wvenc_dbg_table(int which, int i) { init_tables(); return which == 0 ? T_LOG2[i & 255] : T_EXP2[i & 255]; }

static i32 exp2s(i32 log) /* inverse of the decoder's read: WordsUtils.cs:633 */
{
    u32 value;
    if (log < 0) return -exp2s(-log);
    value = T_EXP2[log & 0xff] | 0x100;
    if ((log >>= 8) <= 9) return (i32)(value >> (9 - log));
    return (i32)(value << ((log - 9) & 31));
}
static int mylog2(u32 avalue) /* WordsUtils.cs:588 */
{
    int dbits;
    if ((avalue += avalue >> 9) < (1 << 8)) {
        dbits = bitlen32(avalue);
        return (dbits << 8) + T_LOG2[(avalue << (9 - dbits)) & 0xff];
    }
    dbits = bitlen32(avalue);
    return (dbits << 8) + T_LOG2[(avalue >> (dbits - 9)) & 0xff];
}
static int log2s(i32 v) { return v < 0 ? -mylog2((u32)(-(i64)v)) : mylog2((u32)v); }

static int store_weight(int w) /* inverse of restore_weight, WordsUtils.cs:653 */
{
    if (w > 1024) w = 1024; else if (w < -1024) w = -1024;
    if (w > 0) w -= (w + 64) >> 7;
    return (w + 4) >> 3;
}
static int restore_weight(int8_t w8)
{
    int r = (int)w8 << 3;
    if (r > 0) r += (r + 64) >> 7;
    return r;
}

/* ------------------------------------------------------------------ */
/* PCG32 + signal synthesis (integer only)                              */
/* ------------------------------------------------------------------ */
typedef struct { u64 state, inc; } pcg32;
static inline u32 pcg_next(pcg32 *r)
{
    u64 old = r->state;
    r->state = old * 6364136223846793005ULL + r->inc;
    u32 xs = (u32)(((old >> 18u) ^ old) >> 27u), rot = (u32)(old >> 59u);
    return (xs >> rot) | (xs << ((-rot) & 31));
}
static void pcg_seed(pcg32 *r, u64 seed, u64 seq)
{
    r->state = 0; r->inc = (seq << 1u) | 1u;
    pcg_next(r); r->state += seed; pcg_next(r);
}

typedef struct { i32 x, y, e; int sh; } osc_t; /* "magic circle" oscillator: x -= e*y; y += e*x (Q16 e) */

static void synth_pcm(const wvenc_config *cfg, u64 seed, i64 n, i32 *out, int bits, int nch)
{
    pcg32 rng;
    pcg_seed(&rng, seed, 54u);
    enum { MAXCH = 8, MAXOSC = 5 };
    osc_t osc[MAXCH][MAXOSC];
    int nosc = 3 + (int)(pcg_next(&rng) % 3);
    i32 base_e[MAXOSC];
    int base_sh[MAXOSC];
    for (int k = 0; k < nosc; k++) {
        base_e[k] = 300 + (i32)(pcg_next(&rng) % 24000);
        base_sh[k] = 2 + k + (int)(pcg_next(&rng) % 2); /* amplitude 2^(28-sh) relative to 2^30 full scale */
    }
    for (int c = 0; c < nch; c++)
        for (int k = 0; k < nosc; k++) {
            osc[c][k].e = base_e[k] + (c ? (i32)(pcg_next(&rng) % 64) - 32 : 0);
            osc[c][k].x = (1 << 28) - (i32)(pcg_next(&rng) % (1 << 24));
            osc[c][k].y = (i32)(pcg_next(&rng) % (1 << 26));
            osc[c][k].sh = base_sh[k] + (c && (pcg_next(&rng) & 1) ? 1 : 0);
        }
    /* full scale 2^30 -> target bits: shift right by (31 - bits) after summing (sum < 2^30) */
    int down = 31 - bits;
    u32 noise_mask = bits >= 12 ? ((1u << (bits - 11)) - 1) : 0; /* ~ -60 dBFS triangular */
    i64 gap0 = n > 8 ? (i64)(pcg_next(&rng) % (u32)(n / 2 + 1)) : 0, gaplen = cfg->sample_rate / 4;
    i64 eq0 = n > 8 ? (i64)(n / 2 + pcg_next(&rng) % (u32)(n / 4 + 1)) : 0, eqlen = cfg->sample_rate / 4;
    for (i64 t = 0; t < n; t++) {
        for (int c = 0; c < nch; c++) {
            i64 acc = 0;
            for (int k = 0; k < nosc; k++) {
                osc_t *o = &osc[c][k];
                o->x -= (i32)(((i64)o->e * o->y) >> 16);
                o->y += (i32)(((i64)o->e * o->x) >> 16);
                acc += o->x >> o->sh;
            }
            i32 v = down >= 0 ? (i32)(acc >> down) : (i32)(acc << -down);
            if (noise_mask) v += (i32)(pcg_next(&rng) & noise_mask) - (i32)(pcg_next(&rng) & noise_mask);
            i32 lim = (i32)((1u << (bits - 1)) - 1);
            if (v > lim) v = lim; else if (v < -lim - 1) v = -lim - 1;
            if (t >= gap0 && t < gap0 + gaplen) v = 0;
            out[t * nch + c] = v;
        }
        if (nch >= 2 && t >= eq0 && t < eq0 + eqlen) out[t * nch + 1] = out[t * nch];
    }
}

void wvenc_synth(const wvenc_config *cfg, uint64_t seed, int64_t nsamples, int32_t *out)
{
    int nch = cfg->channels;
    if (cfg->kind == WVENC_DSD) {
        /* 2nd-order sigma-delta of a 16-bit synthetic signal, 8 one-bit samples per output byte, MSB first */
        i32 *pcm = (i32 *)malloc(sizeof(i32) * (size_t)(nsamples + 1) * nch);
        wvenc_config c2 = *cfg;
        c2.kind = WVENC_PCM;
        synth_pcm(&c2, seed, nsamples, pcm, 16, nch); /* one PCM value per byte-time, held for 8 bits */
        for (int c = 0; c < nch; c++) {
            i32 i1 = 0, i2 = 0;
            for (i64 t = 0; t < nsamples; t++) {
                i32 x = pcm[t * nch + c] / 2;
                u32 byte = 0;
                for (int b = 0; b < 8; b++) {
                    i32 q = (i2 >= 0) ? 32767 : -32767;
                    i1 += x - q;
                    i2 += i1 - q;
                    byte = (byte << 1) | (q > 0);
                }
                out[t * nch + c] = (i32)byte;
            }
        }
        free(pcm);
        return;
    }
    int bits = cfg->kind == WVENC_FLOAT ? 24 : cfg->bits;
    synth_pcm(cfg, seed, nsamples, out, bits, nch);
    if (cfg->kind == WVENC_PCM && cfg->bits == 32) {
        /* impose the redundancy the INT32 fields describe (UnpackUtils.cs:1301-1306) */
        i64 total = nsamples * nch;
        for (i64 i = 0; i < total; i++) {
            i32 v = out[i];
            if (cfg->int32_zeros) v = (i32)((u32)v & ~((1u << cfg->int32_zeros) - 1));
            else if (cfg->int32_ones) v = (i32)((u32)v | ((1u << cfg->int32_ones) - 1));
            else if (cfg->int32_dups) {
                u32 m = (1u << cfg->int32_dups) - 1;
                v = ((u32)v >> cfg->int32_dups) & 1 ? (i32)((u32)v | m) : (i32)((u32)v & ~m);
            }
            out[i] = v;
        }
    }
    if (cfg->shift > 0 && cfg->kind == WVENC_PCM) {
        i64 total = nsamples * nch;
        for (i64 i = 0; i < total; i++) out[i] = (i32)((u32)out[i] & ~((1u << cfg->shift) - 1));
    }
}

/* ------------------------------------------------------------------ */
/* bit writer (LSB first, BitsUtils.cs:15-68 read order)                */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t cap, len; u64 acc; int nb; int overflow; } bitw;
static void bw_init(bitw *w, uint8_t *p, size_t cap) { w->p = p; w->cap = cap; w->len = 0; w->acc = 0; w->nb = 0; w->overflow = 0; }
static inline void bw_put(bitw *w, u32 v, int n) /* n <= 32 */
{
    if (!n) return;
    if (n < 32) v &= (1u << n) - 1;
    w->acc |= (u64)v << w->nb;
    w->nb += n;
    while (w->nb >= 8) {
        if (w->len < w->cap) w->p[w->len++] = (uint8_t)w->acc; else w->overflow = 1;
        w->acc >>= 8;
        w->nb -= 8;
    }
}
static inline void bw_ones(bitw *w, int n) { while (n >= 32) { bw_put(w, 0xffffffffu, 32); n -= 32; } if (n) bw_put(w, 0xffffffffu, n); }
static size_t bw_close(bitw *w) /* pad with 1 bits to a byte (the reader sees 0xFF past the end anyway) */
{
    if (w->nb) bw_put(w, 0xffu, 8 - w->nb);
    return w->len;
}
static void bw_gamma(bitw *w, u32 x) /* WordsUtils.cs:321-335 / 391-405 */
{
    if (x < 2) { bw_ones(w, (int)x); bw_put(w, 0, 1); return; }
    int nb = bitlen32(x);
    bw_ones(w, nb);
    bw_put(w, 0, 1);
    bw_put(w, x, nb - 1);
}
static void bw_unary(bitw *w, u32 u) /* WordsUtils.cs:361-414 */
{
    if (u < 16) { bw_ones(w, (int)u); bw_put(w, 0, 1); return; }
    bw_ones(w, 16);
    bw_put(w, 0, 1);
    bw_gamma(w, u - 16);
}

/* ------------------------------------------------------------------ */
/* entropy coder: mirror of get_words (WordsUtils.cs:272-511)           */
/* ------------------------------------------------------------------ */
typedef struct { i32 median[3]; i32 slow_level; i32 error_limit; } ent_t;
typedef struct {
    ent_t c[2];
    int h1, hz;          /* holding_one / holding_zero */
    i64 bitrate_acc[2], bitrate_delta[2];
    u32 flags;
    int mono;            /* MONO_FLAG | FALSE_STEREO */
    /* run pending */
    int run_active; u32 run_count;
    /* word pending (awaiting parity bit of its unary) */
    int pend_active; u32 pend_ubase; u64 pend_code; int pend_codebits; int pend_sign;
    bitw *bw;
} wenc;

#define GET_MED(e, i) (((e)->median[i] >> 4) + 1)
static inline void inc_med(ent_t *e, int i) { static const int D[3] = { 128, 64, 32 }, S[3] = { 7, 6, 5 }; e->median[i] += ((e->median[i] + D[i]) >> S[i]) * 5; }
static inline void dec_med(ent_t *e, int i) { static const int D[3] = { 128, 64, 32 }, S[3] = { 7, 6, 5 }; e->median[i] -= ((e->median[i] + (D[i] - 2)) >> S[i]) * 2; }

static u32 peek_ones(const ent_t *e, i32 value)
{
    u32 a = (u32)(value < 0 ? ~value : value);
    u32 g = (u32)GET_MED(e, 0);
    if (a < g) return 0;
    a -= g;
    g = (u32)GET_MED(e, 1);
    if (a < g) return 1;
    a -= g;
    g = (u32)GET_MED(e, 2);
    return 2 + a / g;
}

static void update_error_limit(wenc *w) /* WordsUtils.cs:195-261 */
{
    i32 b0 = (i32)((w->bitrate_acc[0] += w->bitrate_delta[0]) >> 16);
    if (w->mono) {
        if (w->flags & F_HYB_BITRATE) {
            i32 sl0 = (w->c[0].slow_level + 128) >> 8;
            w->c[0].error_limit = sl0 - b0 > -0x100 ? exp2s(sl0 - b0 + 0x100) : 0;
        } else
            w->c[0].error_limit = exp2s(b0);
    } else {
        i32 b1 = (i32)((w->bitrate_acc[1] += w->bitrate_delta[1]) >> 16);
        if (w->flags & F_HYB_BITRATE) {
            i32 sl0 = (w->c[0].slow_level + 128) >> 8, sl1 = (w->c[1].slow_level + 128) >> 8;
            if (w->flags & F_HYB_BALANCE) {
                i32 balance = (sl1 - sl0 + b1 + 1) >> 1;
                if (balance > b0) { b1 = b0 * 2; b0 = 0; }
                else if (-balance > b0) { b0 = b0 * 2; b1 = 0; }
                else { b1 = b0 + balance; b0 = b0 - balance; }
            }
            w->c[0].error_limit = sl0 - b0 > -0x100 ? exp2s(sl0 - b0 + 0x100) : 0;
            w->c[1].error_limit = sl1 - b1 > -0x100 ? exp2s(sl1 - b1 + 0x100) : 0;
        } else {
            w->c[0].error_limit = exp2s(b0);
            w->c[1].error_limit = exp2s(b1);
        }
    }
}

static void wenc_flush_pend(wenc *w, int parity)
{
    if (!w->pend_active) return;
    bw_unary(w->bw, w->pend_ubase * 2 + (u32)parity);
    if (w->pend_codebits > 32) {
        bw_put(w->bw, (u32)w->pend_code, 32);
        bw_put(w->bw, (u32)(w->pend_code >> 32), w->pend_codebits - 32);
    } else
        bw_put(w->bw, (u32)w->pend_code, w->pend_codebits);
    bw_put(w->bw, (u32)w->pend_sign, 1);
    w->h1 = parity;
    w->hz = !parity;
    w->pend_active = 0;
}

/* Encode one word for channel ch; returns the value the decoder will reconstruct. */
static i32 wenc_word(wenc *w, int ch, i32 target)
{
    ent_t *e = &w->c[ch];
    const int hybrid = (w->flags & F_HYBRID) != 0;

    if (w->pend_active) /* the previous word carried a unary; its parity says whether THIS word's ones_count is non-zero */
        wenc_flush_pend(w, peek_ones(e, target) >= 1);

    if ((w->c[0].median[0] & ~1) == 0 && !w->hz && !w->h1 && (w->c[1].median[0] & ~1) == 0) {
        if (w->run_active) {
            if (target == 0) {
                w->run_count++;
                e->slow_level -= (e->slow_level + 128) >> 8;
                return 0;
            }
            bw_gamma(w->bw, w->run_count);
            w->run_active = 0;
        } else if (target == 0) {
            w->run_active = 1;
            w->run_count = 1;
            e->slow_level -= (e->slow_level + 128) >> 8;
            memset(w->c[0].median, 0, sizeof(w->c[0].median));
            memset(w->c[1].median, 0, sizeof(w->c[1].median));
            return 0;
        } else
            bw_put(w->bw, 0, 1); /* gamma(0): no run */
    }

    u32 a = (u32)(target < 0 ? ~target : target);
    u32 ones = peek_ones(e, target);
    int has_unary;
    if (w->hz) {
        w->hz = 0; /* ones == 0 by construction of the previous parity */
        has_unary = 0;
    } else {
        has_unary = 1;
        w->pend_ubase = ones - (w->h1 ? 1u : 0u);
    }

    if (hybrid && (w->mono || ch == 0)) update_error_limit(w);

    u32 low, high;
    if (ones == 0) {
        low = 0; high = (u32)GET_MED(e, 0) - 1; dec_med(e, 0);
    } else {
        low = (u32)GET_MED(e, 0); inc_med(e, 0);
        if (ones == 1) {
            high = low + (u32)GET_MED(e, 1) - 1; dec_med(e, 1);
        } else {
            low += (u32)GET_MED(e, 1); inc_med(e, 1);
            if (ones == 2) {
                high = low + (u32)GET_MED(e, 2) - 1; dec_med(e, 2);
            } else {
                low += (ones - 2) * (u32)GET_MED(e, 2);
                high = low + (u32)GET_MED(e, 2) - 1; inc_med(e, 2);
            }
        }
    }

    u64 code = 0;
    int codebits = 0;
    u32 mid;
    if (e->error_limit == 0) { /* read_code inverse, WordsUtils.cs:546-570 */
        u32 maxcode = high - low, cv = a - low;
        int bitcount = bitlen32(maxcode);
        mid = a;
        if (bitcount) {
            u32 extras = (u32)((1ull << bitcount) - maxcode - 1);
            if (cv < extras) { code = cv; codebits = bitcount - 1; }
            else { code = ((u64)((cv + extras) >> 1)) | ((u64)((cv + extras) & 1) << (bitcount - 1)); codebits = bitcount; }
        }
    } else { /* bisection, WordsUtils.cs:486-492 */
        mid = (u32)(((u64)high + low + 1) >> 1);
        while (high - low > (u32)e->error_limit) {
            if (a >= mid) { code |= (u64)1 << codebits; low = mid; }
            else high = mid - 1;
            codebits++;
            mid = (u32)(((u64)high + low + 1) >> 1);
        }
    }
    int sign = target < 0;
    i32 decoded = sign ? ~(i32)mid : (i32)mid;

    if (has_unary) {
        w->pend_active = 1;
        w->pend_code = code;
        w->pend_codebits = codebits;
        w->pend_sign = sign;
    } else {
        if (codebits > 32) { bw_put(w->bw, (u32)code, 32); bw_put(w->bw, (u32)(code >> 32), codebits - 32); }
        else bw_put(w->bw, (u32)code, codebits);
        bw_put(w->bw, (u32)sign, 1);
    }
    if (w->flags & F_HYB_BITRATE)
        e->slow_level = e->slow_level - ((e->slow_level + 128) >> 8) + mylog2(mid);
    return decoded;
}

static void wenc_finish(wenc *w)
{
    if (w->pend_active) wenc_flush_pend(w, 0);
    if (w->run_active) { bw_gamma(w->bw, w->run_count); w->run_active = 0; }
}

/* ------------------------------------------------------------------ */
/* decorrelation model (streaming form of UnpackUtils.cs:688-1240)      */
/* ------------------------------------------------------------------ */
typedef struct {
    int term, delta;
    int wA, wB;
    i32 hA[8], hB[8]; /* h[0] most recent output of this pass */
} dpass;

static inline i32 apply_w(int w, i32 s) { return (i32)(((i64)w * s + 512) >> 10); }
static inline void upd(int *w, int delta, i32 s, i32 in) { if (s && in) *w += ((s ^ in) < 0) ? -delta : delta; }
static inline void upd_clip(int *w, int delta, i32 s, i32 in)
{
    if (s && in) {
        if ((s ^ in) < 0) { if ((*w -= delta) < -1024) *w = -1024; }
        else { if ((*w += delta) > 1024) *w = 1024; }
    }
}
static inline i32 pred_pos(const i32 *h, int term)
{
    if (term == 17) return 2 * h[0] - h[1];
    if (term == 18) return (3 * h[0] - h[1]) >> 1;
    return h[term - 1];
}
static inline void push(i32 *h, int term, i32 v)
{
    int n = term > 8 ? 2 : term;
    for (int i = n - 1; i > 0; i--) h[i] = h[i - 1];
    h[0] = v;
}

/* encoder direction: outputs (of the decoder) -> residual inputs, through passes n-1..0 */
static void enc_frame(dpass *p, int n, int stereo, i32 *A, i32 *B)
{
    for (int d = n - 1; d >= 0; d--) {
        dpass *q = &p[d];
        i32 oA = *A, oB = stereo ? *B : 0, iA, iB = 0;
        if (q->term > 0) {
            i32 s = pred_pos(q->hA, q->term);
            iA = oA - apply_w(q->wA, s);
            upd(&q->wA, q->delta, s, iA);
            push(q->hA, q->term, oA);
            if (stereo) {
                s = pred_pos(q->hB, q->term);
                iB = oB - apply_w(q->wB, s);
                upd(&q->wB, q->delta, s, iB);
                push(q->hB, q->term, oB);
            }
        } else if (q->term == -1) {
            iA = oA - apply_w(q->wA, q->hA[0]);
            upd_clip(&q->wA, q->delta, q->hA[0], iA);
            iB = oB - apply_w(q->wB, oA);
            upd_clip(&q->wB, q->delta, oA, iB);
            q->hA[0] = oB;
        } else if (q->term == -2) {
            iB = oB - apply_w(q->wB, q->hB[0]);
            upd_clip(&q->wB, q->delta, q->hB[0], iB);
            iA = oA - apply_w(q->wA, oB);
            upd_clip(&q->wA, q->delta, oB, iA);
            q->hB[0] = oA;
        } else {
            iA = oA - apply_w(q->wA, q->hA[0]);
            upd_clip(&q->wA, q->delta, q->hA[0], iA);
            iB = oB - apply_w(q->wB, q->hB[0]);
            upd_clip(&q->wB, q->delta, q->hB[0], iB);
            q->hB[0] = oA;
            q->hA[0] = oB;
        }
        *A = iA;
        if (stereo) *B = iB;
    }
}

/* decoder direction: residual inputs -> outputs, passes 0..n-1 */
static void dec_frame(dpass *p, int n, int stereo, i32 *A, i32 *B)
{
    for (int d = 0; d < n; d++) {
        dpass *q = &p[d];
        i32 iA = *A, iB = stereo ? *B : 0, oA, oB = 0;
        if (q->term > 0) {
            i32 s = pred_pos(q->hA, q->term);
            oA = iA + apply_w(q->wA, s);
            upd(&q->wA, q->delta, s, iA);
            push(q->hA, q->term, oA);
            if (stereo) {
                s = pred_pos(q->hB, q->term);
                oB = iB + apply_w(q->wB, s);
                upd(&q->wB, q->delta, s, iB);
                push(q->hB, q->term, oB);
            }
        } else if (q->term == -1) {
            oA = iA + apply_w(q->wA, q->hA[0]);
            upd_clip(&q->wA, q->delta, q->hA[0], iA);
            oB = iB + apply_w(q->wB, oA);
            upd_clip(&q->wB, q->delta, oA, iB);
            q->hA[0] = oB;
        } else if (q->term == -2) {
            oB = iB + apply_w(q->wB, q->hB[0]);
            upd_clip(&q->wB, q->delta, q->hB[0], iB);
            oA = iA + apply_w(q->wA, oB);
            upd_clip(&q->wA, q->delta, oB, iA);
            q->hB[0] = oA;
        } else {
            oA = iA + apply_w(q->wA, q->hA[0]);
            upd_clip(&q->wA, q->delta, q->hA[0], iA);
            oB = iB + apply_w(q->wB, q->hB[0]);
            upd_clip(&q->wB, q->delta, q->hB[0], iB);
            q->hB[0] = oA;
            q->hA[0] = oB;
        }
        *A = oA;
        if (stereo) *B = oB;
    }
}

/* ------------------------------------------------------------------ */
/* output assembly                                                      */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t cap, len; int overflow; } obuf;
static void ob_put(obuf *o, const void *src, size_t n)
{
    if (o->len + n > o->cap) { o->overflow = 1; return; }
    memcpy(o->p + o->len, src, n);
    o->len += n;
}
static void ob_u8(obuf *o, unsigned v) { uint8_t b = (uint8_t)v; ob_put(o, &b, 1); }
static void ob_le16(obuf *o, unsigned v) { ob_u8(o, v); ob_u8(o, v >> 8); }
static void ob_le32(obuf *o, u32 v) { ob_le16(o, v); ob_le16(o, v >> 16); }

/* sub-block TLV, MetadataUtils.cs:25-82 */
static void put_meta(obuf *o, unsigned id, const uint8_t *data, size_t len);

/* WavPack 5 block checksum (ID_BLOCK_CHECKSUM, Defines.cs:83; the reference only notes its presence, MetadataUtils.cs:183).
 * Definition as published in WavPack 5's libwavpack (block_add_checksum / WavpackVerifySingleBlock): the header gets
 * HAS_CHECKSUM (0x10000000) and its final ckSize first; then csum = 0xffffffff, csum = csum * 3 + w over every 16-bit
 * little-endian word of the block up to the checksum sub-block; stored as 4 bytes, or as 2 bytes of csum ^ (csum >> 16).
 * It is the last sub-block of the block. */
static void put_block_checksum(obuf *o, size_t hdr_at, int bytes)
{
    if (o->overflow) return;
    u32 flags, cks = (u32)(o->len - hdr_at - 8 + 2 + (size_t)bytes), csum = 0xffffffffu;
    memcpy(&flags, o->p + hdr_at + 24, 4);
    flags |= 0x10000000u;
    memcpy(o->p + hdr_at + 24, &flags, 4);
    memcpy(o->p + hdr_at + 4, &cks, 4);
    for (size_t i = hdr_at; i + 1 < o->len; i += 2) csum = csum * 3u + ((u32)o->p[i] | ((u32)o->p[i + 1] << 8));
    uint8_t tmp[4];
    if (bytes == 2) csum ^= csum >> 16;
    tmp[0] = (uint8_t)csum; tmp[1] = (uint8_t)(csum >> 8); tmp[2] = (uint8_t)(csum >> 16); tmp[3] = (uint8_t)(csum >> 24);
    put_meta(o, 0x2f, tmp, (size_t)bytes);
}

static void put_meta(obuf *o, unsigned id, const uint8_t *data, size_t len)
{
    size_t words = (len + 1) >> 1;
    if (len & 1) id |= 0x40;
    if (words > 255) {
        ob_u8(o, id | 0x80);
        ob_u8(o, (unsigned)words); ob_u8(o, (unsigned)(words >> 8)); ob_u8(o, (unsigned)(words >> 16));
    } else {
        ob_u8(o, id);
        ob_u8(o, (unsigned)words);
    }
    ob_put(o, data, len);
    if (len & 1) ob_u8(o, 0);
}

typedef struct {
    /* per-stream (one per block position within a multichannel segment) carried state */
    int nterms;
    dpass passes[16];
    i32 slow_level[2];
    i32 med[2][3];
    int first;
} stream_state;

typedef struct {
    const wvenc_config *cfg;
    obuf out;
    i64 total_samples;
    u32 total_field;
    int version;
    /* scratch */
    i32 *resid; /* 2*block_samples */
    uint8_t *bits; size_t bits_cap;
    uint8_t *wvx; size_t wvx_cap;
    stream_state ss[4];
} encoder;

static void header_put(obuf *o, u32 cksize, int version, u32 total, i64 block_index, u32 block_samples, u32 flags, u32 crc)
{
    ob_put(o, "wvpk", 4);
    ob_le32(o, cksize);
    ob_le16(o, (unsigned)version);
    ob_u8(o, (unsigned)((block_index >> 32) & 0xff));
    ob_u8(o, 0);
    ob_le32(o, total);
    ob_le32(o, (u32)block_index);
    ob_le32(o, block_samples);
    ob_le32(o, flags);
    ob_le32(o, crc);
}

static int srate_index(int rate)
{
    static const int rates[] = { 6000, 8000, 9600, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 64000, 88200, 96000, 192000 };
    for (int i = 0; i < 15; i++) if (rates[i] == rate) return i;
    return 15;
}

static void wav_header(uint8_t h[44], int nch, int rate, int bytes, u64 data_bytes)
{
    u32 db = (u32)data_bytes;
    memcpy(h, "RIFF", 4);
    u32 v = db + 36; memcpy(h + 4, &v, 4);
    memcpy(h + 8, "WAVEfmt ", 8);
    v = 16; memcpy(h + 16, &v, 4);
    uint16_t s = 1; memcpy(h + 20, &s, 2);
    s = (uint16_t)nch; memcpy(h + 22, &s, 2);
    v = (u32)rate; memcpy(h + 24, &v, 4);
    v = (u32)(rate * nch * bytes); memcpy(h + 28, &v, 4);
    s = (uint16_t)(nch * bytes); memcpy(h + 32, &s, 2);
    s = (uint16_t)(bytes * 8); memcpy(h + 34, &s, 2);
    memcpy(h + 36, "data", 4);
    memcpy(h + 40, &db, 4);
}

/* fixup_samples model used for recon (UnpackUtils.cs:1251-1404, FloatUtils.cs:32-56) */
static i32 shl(i32 v, int n) { return (i32)((u32)v << (n & 31)); }
static i32 model_fixup_one(const wvenc_config *cfg, u32 flags, i32 v, int has_wvx, u32 wvx_data, int bits_read)
{
    int shift = (int)((flags >> 13) & 0x1f);
    int lossy = (flags & F_HYBRID) != 0;
    if (flags & F_FLOAT) {
        int s = cfg->float_max_exp - cfg->float_norm_exp + cfg->float_shift;
        if (s > 32) s = 32; else if (s < -32) s = -32;
        if (s > 0) v = shl(v, s); else if (s < 0) v = v >> ((-s) & 31);
        if (v > 8388607) v = 8388607; else if (v < -8388608) v = -8388608;
        return v;
    }
    if (flags & F_INT32) {
        int sent = cfg->int32_sent_bits, zeros = cfg->int32_zeros, ones = cfg->int32_ones, dups = cfg->int32_dups;
        if (has_wvx) {
            if (sent > 0) {
                if (bits_read >= 0) v = shl((i32)((u32)shl(v, bits_read) | wvx_data), sent - bits_read);
                else v = shl(v, sent);
            }
            if (zeros) v = shl(v, zeros);
            else if (ones) v = shl(v + 1, ones) - 1;
            else if (dups) v = shl(v + (v & 1), dups) - (v & 1);
        } else if (sent == 0 && (zeros + ones + dups) != 0) {
            while (lossy && (flags & 3) == 3 && shift < 8) {
                if (zeros > 0) zeros--; else if (ones > 0) ones--; else if (dups > 0) dups--; else break;
                shift++;
            }
            if (zeros) v = shl(v, zeros);
            else if (ones) v = shl(v + 1, ones) - 1;
            else if (dups) v = shl(v + (v & 1), dups) - (v & 1);
        } else
            shift += zeros + sent + ones + dups;
    }
    shift &= 0x1f;
    if (lossy) {
        i32 mn, mx;
        switch (flags & 3) {
        case 0: mn = -128 >> shift; mx = 127 >> shift; break;
        case 1: mn = -32768 >> shift; mx = 32767 >> shift; break;
        case 2: mn = -8388608 >> shift; mx = 8388607 >> shift; break;
        default: mn = (i32)(0x80000000u >> shift); mx = 0x7FFFFFFF >> shift; break;
        }
        if (v < mn) v = shl(mn, shift); else if (v > mx) v = shl(mx, shift); else v = shl(v, shift);
    } else if (shift)
        v = shl(v, shift);
    return v;
}

/* Encode one mono/stereo PCM-family block.  src: nch_blk-interleaved source samples (n frames). */
static void encode_pcm_block(encoder *E, stream_state *S, const i32 *srcL, const i32 *srcR, int stride, i64 n, i64 block_index,
                             u32 pos_flags, int first_block_of_file, int total_channels, i32 *reconL, i32 *reconR)
{
    const wvenc_config *cfg = E->cfg;
    int stereo_src = srcR != NULL;
    int false_stereo = 0;
    if (stereo_src && cfg->false_stereo) {
        false_stereo = 1;
        for (i64 t = 0; t < n; t++) if (srcL[t * stride] != srcR[t * stride]) { false_stereo = 0; break; }
    }
    if (cfg->kind == WVENC_PCM && cfg->bits == 32 && cfg->int32_sent_bits && cfg->int32_wvx)
        false_stereo = 0; /* quirk C-6: the reference would pull WVX bits for 2n values */
    int stereo = stereo_src && !false_stereo; /* coded as two channels */
    int bytes = cfg->kind == WVENC_FLOAT ? 4 : (cfg->bits + 7) / 8;
    u32 flags = (u32)(bytes - 1) | pos_flags;
    if (!stereo_src) flags |= F_MONO;
    if (false_stereo) flags |= F_FALSE_STEREO;
    if (stereo && cfg->joint_stereo) flags |= F_JOINT;
    int hybrid = cfg->kind == WVENC_HYBRID;
    if (hybrid) flags |= F_HYBRID | F_HYB_BITRATE | (cfg->hybrid_balance && stereo ? F_HYB_BALANCE : 0);
    if (cfg->kind == WVENC_FLOAT) flags |= F_FLOAT;
    int int32_mode = cfg->kind == WVENC_PCM && cfg->bits == 32 &&
                     (cfg->int32_sent_bits || cfg->int32_zeros || cfg->int32_ones || cfg->int32_dups);
    if (int32_mode) flags |= F_INT32;
    int shift = cfg->kind == WVENC_PCM && !int32_mode ? cfg->shift : 0;
    flags |= (u32)shift << 13;
    int sri = (cfg->extras & WVENC_X_SAMPLE_RATE) ? 15 : srate_index(cfg->sample_rate);
    flags |= (u32)sri << 23;

    /* pass list for this block (decoder array order = reverse file order); mono blocks drop cross-channel terms */
    int nt = 0;
    int8_t fterm[16], fdelta[16];
    for (int i = 0; i < cfg->nterms; i++) {
        if (!stereo && cfg->terms[i] < 0) continue;
        fterm[nt] = cfg->terms[i]; fdelta[nt] = cfg->deltas[i]; nt++;
    }
    if (S->first || S->nterms != nt) {
        memset(S->passes, 0, sizeof(S->passes));
        S->nterms = nt;
        S->slow_level[0] = S->slow_level[1] = 0;
        memset(S->med, 0, sizeof(S->med));
    }
    dpass *P = S->passes; /* decoder order: P[d] <-> file index nt-1-d */
    int8_t w8A[16], w8B[16];
    for (int d = 0; d < nt; d++) {
        int f = nt - 1 - d;
        int term_changed = P[d].term != fterm[f];
        P[d].term = fterm[f]; P[d].delta = fdelta[f];
        if (term_changed) { P[d].wA = P[d].wB = 0; memset(P[d].hA, 0, sizeof(P[d].hA)); memset(P[d].hB, 0, sizeof(P[d].hB)); }
        w8A[d] = (int8_t)store_weight(P[d].wA); w8B[d] = (int8_t)store_weight(P[d].wB);
        P[d].wA = restore_weight(w8A[d]);
        P[d].wB = stereo ? restore_weight(w8B[d]) : 0;
        if (!stereo) { memset(P[d].hB, 0, sizeof(P[d].hB)); }
    }
    /* history: stored through 16-bit logs; only the encoder-first pass (decoder's last) unless ALL_HISTORY */
    int16_t hlog[16][2][8];
    for (int d = 0; d < nt; d++) {
        int keep = (d == nt - 1) || (cfg->extras & WVENC_X_ALL_HISTORY);
        int cnt = P[d].term > 8 ? 2 : (P[d].term < 0 ? 1 : P[d].term);
        for (int j = 0; j < 8; j++) {
            if (!keep || j >= cnt) { P[d].hA[j] = 0; P[d].hB[j] = 0; if (j < 8) { hlog[d][0][j] = hlog[d][1][j] = 0; } continue; }
            hlog[d][0][j] = (int16_t)log2s(P[d].hA[j]); P[d].hA[j] = exp2s(hlog[d][0][j]);
            hlog[d][1][j] = (int16_t)log2s(P[d].hB[j]); P[d].hB[j] = exp2s(hlog[d][1][j]);
        }
        if (!stereo) for (int j = 0; j < 8; j++) { P[d].hB[j] = 0; hlog[d][1][j] = 0; }
        if (P[d].term == -1) { P[d].hB[0] = 0; }   /* only samples_A[0] is live for -1, samples_B[0] for -2 */
        if (P[d].term == -2) { P[d].hA[0] = 0; }
    }

    /* entropy state */
    bitw bw;
    bw_init(&bw, E->bits, E->bits_cap);
    wenc W;
    memset(&W, 0, sizeof(W));
    W.bw = &bw;
    W.flags = flags;
    W.mono = !stereo;
    uint8_t ent_bytes[12];
    {
        /* initial medians = previous block's final medians, stored as unsigned 16-bit logs (WordsUtils.cs:98-110) */
        for (int c = 0; c < (stereo ? 2 : 1); c++)
            for (int k = 0; k < 3; k++) {
                int l = S->first ? 0 : mylog2((u32)S->med[c][k]);
                ent_bytes[c * 6 + k * 2] = (uint8_t)l; ent_bytes[c * 6 + k * 2 + 1] = (uint8_t)(l >> 8);
                W.c[c].median[k] = exp2s(l);
            }
    }
    uint8_t hyb_bytes[12];
    int hyb_len = 0;
    if (hybrid) {
        for (int c = 0; c < (stereo ? 2 : 1); c++) {
            int l = mylog2((u32)S->slow_level[c]);
            hyb_bytes[hyb_len++] = (uint8_t)l; hyb_bytes[hyb_len++] = (uint8_t)(l >> 8);
            W.c[c].slow_level = exp2s(l);
        }
        for (int c = 0; c < (stereo ? 2 : 1); c++) {
            int br = cfg->hybrid_bitrate & 0xffff;
            hyb_bytes[hyb_len++] = (uint8_t)br; hyb_bytes[hyb_len++] = (uint8_t)(br >> 8);
            W.bitrate_acc[c] = (i64)(i32)((u32)br << 16);
        }
        if (block_index & 1) { /* every other block carries a (small) bitrate_delta to exercise that field */
            for (int c = 0; c < (stereo ? 2 : 1); c++) {
                int dl = log2s(c ? -3 : 5);
                hyb_bytes[hyb_len++] = (uint8_t)dl; hyb_bytes[hyb_len++] = (uint8_t)(dl >> 8);
                W.bitrate_delta[c] = exp2s((int16_t)dl);
            }
        }
    }

    /* INT32 / shift pre-transform of the source into the coded integer domain */
    int sent = int32_mode ? cfg->int32_sent_bits : 0;
    int zeros = int32_mode ? cfg->int32_zeros : 0, ones_f = int32_mode ? cfg->int32_ones : 0, dups = int32_mode ? cfg->int32_dups : 0;
    int has_wvx = int32_mode && sent > 0 && cfg->int32_wvx;
    int max_width = has_wvx && cfg->int32_new_wvx ? (cfg->int32_max_width & 31) : 0;
    bitw xw;
    bw_init(&xw, E->wvx, E->wvx_cap);
    if (has_wvx && cfg->int32_new_wvx) bw_put(&xw, (u32)max_width, 5);
    if (cfg->kind == WVENC_FLOAT && cfg->float_new_wvx) { bw_put(&xw, 3, 5); bw_put(&xw, 7, 5); }

    u32 crc = 0xffffffffu, crc_x = 0xffffffffu, magor = 0;
    int nchb = stereo_src ? 2 : 1;

    for (i64 t = 0; t < n; t++) {
        i32 y[2], v[2];
        y[0] = srcL[t * stride];
        y[1] = stereo_src ? srcR[t * stride] : 0;
        for (int c = 0; c < nchb; c++) {
            i32 x = y[c];
            if (int32_mode) {
                if (zeros) x >>= zeros; else if (ones_f) x >>= ones_f; else if (dups) x >>= dups;
                y[c] = x;             /* value after undoing redundancy: (v << sent) | data */
                v[c] = x >> sent;
            } else
                v[c] = x >> shift;
        }
        i32 A = v[0], B = v[1];
        if (stereo && (flags & F_JOINT)) { A = v[0] - v[1]; B = v[1] + (A >> 1); } /* inverse of UnpackUtils.cs:615 */
        i32 oA, oB;
        if (!hybrid) {
            i32 rA = A, rB = B;
            enc_frame(P, nt, stereo, &rA, &rB);
            wenc_word(&W, 0, rA);
            if (stereo) wenc_word(&W, 1, rB);
            oA = A; oB = B;
        } else {
            /* closed loop: choose residuals against a trial decode, then commit the decoder step */
            dpass T[16];
            i32 a0 = 0, b0 = 0;
            memcpy(T, P, sizeof(dpass) * (size_t)nt);
            dec_frame(T, nt, stereo, &a0, &b0);
            i32 ra = wenc_word(&W, 0, A - a0), rb = 0;
            if (stereo) {
                i32 a1 = ra, b1 = 0;
                memcpy(T, P, sizeof(dpass) * (size_t)nt);
                dec_frame(T, nt, stereo, &a1, &b1);
                rb = wenc_word(&W, 1, B - b1);
            }
            oA = ra; oB = rb;
            dec_frame(P, nt, stereo, &oA, &oB);
        }
        /* decoder-side view after decorrelation: (oA,oB) -> joint undo -> values that feed CRC and fixup */
        i32 dv[2] = { oA, oB };
        if (stereo && (flags & F_JOINT)) { dv[1] = oB - (oA >> 1); dv[0] = oA + dv[1]; }
        for (int c = 0; c < (stereo ? 2 : 1); c++) {
            crc = crc * 3 + (u32)dv[c];
            magor |= (u32)(dv[c] < 0 ? ~dv[c] : dv[c]);
        }
        /* WVX bits + recon */
        for (int c = 0; c < nchb; c++) {
            int cc = stereo ? c : 0;
            i32 val = dv[cc];
            u32 data = 0;
            int btr = -1;
            if (has_wvx && !(c == 1 && false_stereo)) {
                u32 mask = sent >= 32 ? 0xffffffffu : ((1u << sent) - 1);
                u32 low = (u32)y[c] & mask;
                btr = sent;
                if (max_width > 0) {
                    i32 pv = val < 0 ? ~val : val;
                    int width = bitlen32((u32)pv) + sent;
                    if (!(width <= max_width || (btr -= width - max_width) > 0)) btr = -1;
                }
                if (btr >= 0) { data = low >> (sent - btr); bw_put(&xw, data, btr); }
            }
            i32 r = model_fixup_one(cfg, flags, val, has_wvx, data, btr);
            if (has_wvx && !(c == 1 && false_stereo)) crc_x = crc_x * 9 + ((u32)r & 0xffff) * 3 + (((u32)r >> 16) & 0xffff);
            if (c == 0 && reconL) reconL[t * stride] = r;
            if (c == 1 && reconR) reconR[t * stride] = false_stereo ? reconL[t * stride] : r;
        }
    }
    /* quirk C-6: FALSE_STEREO + INT32 + WVX makes the decoder pull WVX bits for 2n values; keep that case out of the corpus */
    wenc_finish(&W);
    size_t wvlen = bw_close(&bw);
    if (wvlen & 1) { if (wvlen < E->bits_cap) E->bits[wvlen++] = 0xff; }
    if (wvlen == 0) { E->bits[0] = E->bits[1] = 0xff; wvlen = 2; }
    if (hybrid) { S->slow_level[0] = W.c[0].slow_level; S->slow_level[1] = W.c[1].slow_level; }
    for (int c = 0; c < 2; c++) for (int k = 0; k < 3; k++) S->med[c][k] = W.c[c].median[k];
    S->first = 0;

    int mag = bitlen32(magor);
    if (mag > 31) mag = 31;
    flags |= (u32)mag << 18;

    /* ---- assemble ---- */
    obuf *o = &E->out;
    size_t hdr_at = o->len;
    header_put(o, 0, E->version, E->total_field, block_index, (u32)n, flags, crc);
    uint8_t tmp[256];
    if (cfg->extras & WVENC_X_DUMMY) { put_meta(o, 0x00, (const uint8_t *)"\0\0", 2); put_meta(o, 0x3e, (const uint8_t *)"odd", 3); }
    /* terms (file order) */
    for (int f = 0; f < nt; f++) tmp[f] = (uint8_t)(((fterm[f] + 5) & 0x1f) | ((fdelta[f] & 7) << 5));
    put_meta(o, 0x02, tmp, (size_t)nt);
    { /* weights, file order = decoder index nt-1 downwards */
        int k = 0;
        for (int d = nt - 1; d >= 0; d--) { tmp[k++] = (uint8_t)w8A[d]; if (stereo) tmp[k++] = (uint8_t)w8B[d]; }
        put_meta(o, 0x03, tmp, (size_t)k);
    }
    { /* history */
        int k = 0;
        if (E->version == 0x402 && hybrid) { /* v0x402 hybrid streams carry 2 (mono) / 4 (stereo) extra bytes first (UnpackUtils.cs:277-283) */
            int pad = stereo ? 4 : 2;
            for (int q = 0; q < pad; q++) tmp[k++] = 0;
        }
        int npass = (cfg->extras & WVENC_X_ALL_HISTORY) ? nt : (nt ? 1 : 0);
        for (int q = 0; q < npass; q++) {
            int d = nt - 1 - q;
            int term = P[d].term;
            if (term > 8) {
                for (int j = 0; j < 2; j++) { tmp[k++] = (uint8_t)hlog[d][0][j]; tmp[k++] = (uint8_t)(hlog[d][0][j] >> 8); }
                if (stereo) for (int j = 0; j < 2; j++) { tmp[k++] = (uint8_t)hlog[d][1][j]; tmp[k++] = (uint8_t)(hlog[d][1][j] >> 8); }
            } else if (term < 0) {
                tmp[k++] = (uint8_t)hlog[d][0][0]; tmp[k++] = (uint8_t)(hlog[d][0][0] >> 8);
                tmp[k++] = (uint8_t)hlog[d][1][0]; tmp[k++] = (uint8_t)(hlog[d][1][0] >> 8);
            } else {
                /* decoder slot m holds x[-term+m] (UnpackUtils.cs:888-918); our h[j] is x[-1-j] */
                for (int m = 0; m < term; m++) {
                    int j = term - 1 - m;
                    tmp[k++] = (uint8_t)hlog[d][0][j]; tmp[k++] = (uint8_t)(hlog[d][0][j] >> 8);
                    if (stereo) { tmp[k++] = (uint8_t)hlog[d][1][j]; tmp[k++] = (uint8_t)(hlog[d][1][j] >> 8); }
                }
            }
        }
        put_meta(o, 0x04, tmp, (size_t)k);
    }
    put_meta(o, 0x05, ent_bytes, stereo ? 12 : 6);
    if (hybrid) put_meta(o, 0x06, hyb_bytes, (size_t)hyb_len);
    if (cfg->kind == WVENC_FLOAT) {
        tmp[0] = (uint8_t)cfg->float_flags; tmp[1] = (uint8_t)cfg->float_shift; tmp[2] = (uint8_t)cfg->float_max_exp; tmp[3] = (uint8_t)cfg->float_norm_exp;
        put_meta(o, 0x08, tmp, 4);
    }
    if (int32_mode) {
        tmp[0] = (uint8_t)sent; tmp[1] = (uint8_t)zeros; tmp[2] = (uint8_t)ones_f; tmp[3] = (uint8_t)dups;
        put_meta(o, 0x09, tmp, 4);
    }
    if (total_channels > 2) {
        tmp[0] = (uint8_t)total_channels; tmp[1] = 0x3f;
        put_meta(o, 0x0d, tmp, 2);
    }
    if (first_block_of_file) {
        if (cfg->extras & WVENC_X_RIFF_HEADER) {
            uint8_t h[44];
            int ob = cfg->kind == WVENC_FLOAT ? 4 : bytes;
            wav_header(h, total_channels, cfg->sample_rate, ob, (u64)E->total_samples * total_channels * ob);
            put_meta(o, 0x21, h, 44);
        }
        if (cfg->extras & WVENC_X_CONFIG) {
            u32 cf = (hybrid ? 8u : 0u) | (stereo && cfg->joint_stereo ? 0x10u : 0u) | (cfg->kind == WVENC_FLOAT ? 0x80u : 0u);
            if (cfg->nterms >= 10) cf |= 0x800;   /* HIGH */
            if (cfg->nterms >= 16) cf |= 0x1000;  /* VERY_HIGH */
            if (cfg->nterms <= 2) cf |= 0x200;    /* FAST */
            if (cfg->extras & WVENC_X_MD5_TRAILER) cf |= 0x8000000;
            if (cfg->extras & WVENC_X_ALL_HISTORY) cf |= 0x2000000; /* pretend EXTRA mode so xmode is exercised */
            tmp[0] = (uint8_t)(cf >> 8); tmp[1] = (uint8_t)(cf >> 16); tmp[2] = (uint8_t)(cf >> 24); tmp[3] = 3;
            put_meta(o, 0x25, tmp, (cf & 0x2000000) ? 4 : 3);
        }
        if (cfg->extras & WVENC_X_NEW_CONFIG) { tmp[0] = 0; put_meta(o, 0x2a, tmp, 1); }
    }
    if (cfg->extras & WVENC_X_SAMPLE_RATE) { /* every block: blocks are self-describing */
        tmp[0] = (uint8_t)cfg->sample_rate; tmp[1] = (uint8_t)(cfg->sample_rate >> 8); tmp[2] = (uint8_t)(cfg->sample_rate >> 16);
        put_meta(o, 0x27, tmp, 3);
    }
    put_meta(o, 0x0a, E->bits, wvlen);
    if (has_wvx || cfg->kind == WVENC_FLOAT) {
        size_t xl = bw_close(&xw);
        /* sub-block = 4-byte crc + payload, even length, > 4 bytes */
        size_t plen = 4 + xl;
        if (plen < 6) plen = 6;
        if (plen & 1) plen++;
        uint8_t *blob = (uint8_t *)malloc(plen);
        memset(blob, 0xff, plen);
        blob[0] = (uint8_t)crc_x; blob[1] = (uint8_t)(crc_x >> 8); blob[2] = (uint8_t)(crc_x >> 16); blob[3] = (uint8_t)(crc_x >> 24);
        memcpy(blob + 4, E->wvx, xl);
        int newfmt = cfg->kind == WVENC_FLOAT ? cfg->float_new_wvx : cfg->int32_new_wvx;
        put_meta(o, newfmt ? 0x2c : 0x0c, blob, plen);
        free(blob);
    }
    if (cfg->extras & WVENC_X_BLOCK_CHECKSUM) put_block_checksum(o, hdr_at, 4);
    if (!o->overflow) {
        u32 cks = (u32)(o->len - hdr_at - 8);
        memcpy(o->p + hdr_at + 4, &cks, 4);
    }
    if (bw.overflow || xw.overflow) o->overflow = 1;
}

/* ------------------------------------------------------------------ */
/* DSD blocks (DsdUtils.cs)                                             */
/* ------------------------------------------------------------------ */
typedef struct { uint8_t *p; size_t cap, len; u32 low, high; int overflow; } rcw;
static inline void rc_emit(rcw *r) { if (r->len < r->cap) r->p[r->len++] = (uint8_t)(r->high >> 24); else r->overflow = 1; }
static inline void rc_norm(rcw *r)
{
    while (((r->high ^ r->low) & 0xFF000000u) == 0) { rc_emit(r); r->high = (r->high << 8) | 0xFF; r->low <<= 8; }
}
static inline void rc_flush4(rcw *r) /* emit the 4 bytes of low; leaves low=0, high=0xFFFFFFFF */
{
    r->high = r->low;
    for (int i = 0; i < 4; i++) { rc_emit(r); r->high = (r->high << 8) | 0xFF; r->low <<= 8; }
}

static void init_ptable(i32 *table, int rate_i, int rate_s) /* DsdUtils.cs:321-341 */
{
    i32 value = 0x808000, rate = rate_i << 8, c, i;
    for (c = (rate + 128) >> 8; c > 0; c--) value += (0x10000 - value) >> 8;
    for (i = 0; i < 128; ++i) {
        table[i] = value;
        table[255 - i] = 0x100ffff - value;
        if (value > 0x010000) {
            rate += (rate * rate_s + 128) >> 8;
            for (c = (rate + 64) >> 7; c > 0; c--) value += (0x10000 - value) >> 8;
        }
    }
}

typedef struct { i32 value, filter0, filter1, filter2, filter3, filter4, filter5, filter6, factor; } dsdf;

static size_t encode_dsd_payload(const wvenc_config *cfg, const i32 *src, int nch, i64 n, uint8_t *out, size_t cap)
{
    size_t k = 0;
    i64 total = n * nch;
    if (cap < 64) return 0;
    out[k++] = (uint8_t)cfg->dsd_rate_shift;
    out[k++] = (uint8_t)cfg->dsd_mode;
    if (cfg->dsd_mode == 0) {
        if (cap < k + (size_t)total) return 0;
        for (i64 i = 0; i < total; i++) out[k++] = (uint8_t)src[i];
        return k;
    }
    if (cfg->dsd_mode == 1) { /* DsdUtils.cs:149-304 */
        int hb = cfg->dsd_history_bits, bins = 1 << hb;
        u32 *hist = (u32 *)calloc((size_t)bins * 256, sizeof(u32));
        uint8_t *prob = (uint8_t *)calloc((size_t)bins * 256, 1);
        uint16_t *summed = (uint16_t *)calloc((size_t)bins * 256, sizeof(uint16_t));
        int p0 = 0, p1 = 0;
        for (i64 i = 0; i < total; i++) {
            int code = src[i] & 0xff;
            hist[p0 * 256 + code]++;
            if (nch == 1) p0 = code & (bins - 1); else { p0 = p1; p1 = code & (bins - 1); }
        }
        int maxp = 0;
        int cap_p = cfg->dsd_raw_probs ? 255 : 200;
        for (int b = 0; b < bins; b++) {
            u64 tot = 0;
            for (int i = 0; i < 256; i++) tot += hist[b * 256 + i];
            unsigned sum = 0;
            for (int i = 0; i < 256; i++) {
                u32 h = hist[b * 256 + i];
                unsigned p = 0;
                if (h) { p = (unsigned)(((u64)h * 1024 + tot - 1) / tot); if (p > (unsigned)cap_p) p = (unsigned)cap_p; if (!p) p = 1; }
                prob[b * 256 + i] = (uint8_t)p;
                sum += p;
                summed[b * 256 + i] = (uint16_t)sum;
                if ((int)p > maxp) maxp = (int)p;
            }
        }
        out[k++] = (uint8_t)hb;
        if (cfg->dsd_raw_probs) {
            out[k++] = 0xFF;
            if (cap < k + (size_t)bins * 256 + 16) { k = 0; goto done1; }
            memcpy(out + k, prob, (size_t)bins * 256);
            k += (size_t)bins * 256;
        } else {
            if (maxp < 1) maxp = 1;
            out[k++] = (uint8_t)maxp;
            int tot = bins * 256, i = 0;
            while (i < tot) {
                if (k + 8 > cap) { k = 0; goto done1; }
                if (prob[i]) { out[k++] = prob[i++]; continue; }
                int z = 0;
                while (i + z < tot && !prob[i + z] && z < 255 - maxp) z++;
                out[k++] = (uint8_t)(maxp + z);
                i += z;
            }
            out[k++] = 0; /* terminator consumed by DsdUtils.cs:193 */
        }
        {
            rcw r = { out + k, cap - k, 0, 0, 0xFFFFFFFFu, 0 };
            p0 = p1 = 0;
            for (i64 i = 0; i < total; i++) {
                int code = src[i] & 0xff;
                u32 sum = summed[p0 * 256 + 255];
                u32 mult = (r.high - r.low) / sum;
                if (mult == 0) { rc_flush4(&r); mult = r.high / sum; }
                if (code > 0) r.low += summed[p0 * 256 + code - 1] * mult;
                r.high = r.low + prob[p0 * 256 + code] * mult - 1;
                if (nch == 1) p0 = code & (bins - 1); else { p0 = p1; p1 = code & (bins - 1); }
                rc_norm(&r);
            }
            rc_flush4(&r);
            k = r.overflow ? 0 : k + r.len;
        }
    done1:
        free(hist); free(prob); free(summed);
        return k;
    }
    if (cfg->dsd_mode == 3) { /* DsdUtils.cs:343-493 */
        i32 ptable[256];
        dsdf sp[2];
        memset(sp, 0, sizeof(sp));
        int rate_i = cfg->dsd_rate_i, rate_s = 20;
        out[k++] = (uint8_t)rate_i;
        out[k++] = (uint8_t)rate_s;
        init_ptable(ptable, rate_i, rate_s);
        for (int c = 0; c < nch; c++) {
            static const uint8_t finit[5] = { 0x80, 0x80, 0x80, 0x80, 0x80 };
            i32 *f = &sp[c].filter1;
            for (int j = 0; j < 5; j++) { out[k++] = finit[j]; f[j] = finit[j] << 12; }
            sp[c].filter6 = 0;
            int16_t fac = (int16_t)(c ? -40 : 25);
            out[k++] = (uint8_t)fac; out[k++] = (uint8_t)((uint16_t)fac >> 8);
            sp[c].factor = fac;
        }
        rcw r = { out + k, cap - k, 0, 0, 0xFFFFFFFFu, 0 };
        for (i64 t = 0; t < n; t++) {
            for (int c = 0; c < nch; c++) sp[c].value = sp[c].filter1 - sp[c].filter5 + ((sp[c].filter6 * sp[c].factor) >> 2);
            for (int b = 7; b >= 0; b--) {
                for (int c = 0; c < nch; c++) {
                    dsdf *s = &sp[c];
                    int bit = (src[t * nch + c] >> b) & 1;
                    int pp = (s->value >> 8) & 255;
                    u32 split = r.low + ((r.high - r.low) >> 8) * ((u32)ptable[pp] >> 16);
                    if (bit) { r.high = split; ptable[pp] += (0x010000FE - ptable[pp]) >> 8; s->filter0 = -1; }
                    else { r.low = split + 1; ptable[pp] += (0x00010000 - ptable[pp]) >> 8; s->filter0 = 0; }
                    rc_norm(&r);
                    s->value += s->filter6 * 8;
                    s->factor += (((s->value ^ s->filter0) >> 31) | 1) & ((s->value ^ (s->value - (s->filter6 * 16))) >> 31);
                    s->filter1 += ((s->filter0 & (1 << 20)) - s->filter1) >> 6;
                    s->filter2 += ((s->filter0 & (1 << 20)) - s->filter2) >> 4;
                    s->filter3 += (s->filter2 - s->filter3) >> 4;
                    s->filter4 += (s->filter3 - s->filter4) >> 4;
                    s->value = (s->filter4 - s->filter5) >> 4;
                    s->filter5 += s->value;
                    s->filter6 += (s->value - s->filter6) >> 3;
                    s->value = s->filter1 - s->filter5 + ((s->filter6 * s->factor) >> 2);
                }
            }
            for (int c = 0; c < nch; c++) sp[c].factor -= (sp[c].factor + 512) >> 10;
        }
        rc_flush4(&r);
        return r.overflow ? 0 : k + r.len;
    }
    return 0;
}

static void encode_dsd_block(encoder *E, const i32 *src, int nch, i64 n, i64 block_index, int first_block)
{
    const wvenc_config *cfg = E->cfg;
    obuf *o = &E->out;
    u32 flags = F_DSD | F_INITIAL | F_FINAL | (nch == 1 ? F_MONO : 0) | ((u32)srate_index(cfg->sample_rate) << 23);
    u32 crc = 0xffffffffu;
    for (i64 i = 0; i < n * nch; i++) crc = crc * 3 + (u32)(src[i] & 0xff);
    size_t hdr_at = o->len;
    header_put(o, 0, E->version, E->total_field, block_index, (u32)n, flags, crc);
    uint8_t tmp[8];
    if (first_block && (cfg->extras & WVENC_X_CONFIG)) { tmp[0] = 0; tmp[1] = 0; tmp[2] = 0; put_meta(o, 0x25, tmp, 3); }
    if (first_block && (cfg->extras & WVENC_X_NEW_CONFIG)) { tmp[0] = 4; put_meta(o, 0x2a, tmp, 1); } /* DSF */
    size_t cap = (size_t)(n * nch) * 2 + 16384;
    uint8_t *pl = (uint8_t *)malloc(cap);
    size_t len = encode_dsd_payload(cfg, src, nch, n, pl, cap);
    if (!len) o->overflow = 1;
    else put_meta(o, 0x0e, pl, len);
    free(pl);
    if (cfg->extras & WVENC_X_BLOCK_CHECKSUM) put_block_checksum(o, hdr_at, 2);
    if (!o->overflow) {
        u32 cks = (u32)(o->len - hdr_at - 8);
        memcpy(o->p + hdr_at + 4, &cks, 4);
    }
}

/* ------------------------------------------------------------------ */
/* file level                                                           */
/* ------------------------------------------------------------------ */
void wvenc_default_config(wvenc_config *cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->kind = WVENC_PCM;
    cfg->sample_rate = 44100;
    cfg->bits = 16;
    cfg->channels = 2;
    cfg->block_samples = 22050;
    static const int8_t t[5] = { 18, 18, 2, 3, -2 };
    cfg->nterms = 5;
    for (int i = 0; i < 5; i++) { cfg->terms[i] = t[i]; cfg->deltas[i] = 2; }
    cfg->joint_stereo = 1;
    cfg->int32_wvx = 1;
    cfg->hybrid_bitrate = 4 * 256;
    cfg->float_norm_exp = 127; cfg->float_max_exp = 127;
    cfg->dsd_rate_shift = 3; cfg->dsd_history_bits = 4; cfg->dsd_rate_i = 32;
    cfg->extras = WVENC_X_RIFF_HEADER | WVENC_X_CONFIG;
    cfg->version = 0x410;
}

size_t wvenc_bound(const wvenc_config *cfg, int64_t nsamples)
{
    i64 nblocks = cfg->block_samples > 0 ? (nsamples + cfg->block_samples - 1) / cfg->block_samples + 2 : 2;
    int streams = cfg->channels > 2 ? 4 : 1;
    return (size_t)(nsamples * cfg->channels * 6 + nblocks * streams * 1024 + 65536 * streams + (cfg->kind == WVENC_DSD ? nblocks * 32768 : 0));
}

size_t wvenc_encode(const wvenc_config *cfg, const int32_t *samples, int64_t nsamples, uint8_t *out, size_t cap, int32_t *recon)
{
    init_tables();
    if (cfg->block_samples <= 0 || cfg->nterms < 0 || cfg->nterms > 16) return 0;
    if (cfg->channels != 1 && cfg->channels != 2 && cfg->channels != 6) return 0;
    encoder E;
    memset(&E, 0, sizeof(E));
    E.cfg = cfg;
    E.out.p = out; E.out.cap = cap;
    E.total_samples = nsamples;
    E.total_field = cfg->unknown_length ? 0xFFFFFFFFu : (u32)nsamples;
    E.version = cfg->version ? cfg->version : 0x410;
    E.bits_cap = (size_t)cfg->block_samples * 2 * 10 + 4096;
    E.bits = (uint8_t *)malloc(E.bits_cap);
    E.wvx_cap = (size_t)cfg->block_samples * 2 * 4 + 4096;
    E.wvx = (uint8_t *)malloc(E.wvx_cap);
    for (int i = 0; i < 4; i++) E.ss[i].first = 1;
    int nch = cfg->channels;
    i64 done = 0;
    int first = 1;
    while (done < nsamples) {
        i64 n = nsamples - done < cfg->block_samples ? nsamples - done : cfg->block_samples;
        const i32 *base = samples + done * nch;
        i32 *rbase = recon ? recon + done * nch : NULL;
        if (cfg->kind == WVENC_DSD) {
            encode_dsd_block(&E, base, nch, n, done, first);
            if (recon) memcpy(rbase, base, sizeof(i32) * (size_t)(n * nch));
        } else if (nch == 1)
            encode_pcm_block(&E, &E.ss[0], base, NULL, 1, n, done, F_INITIAL | F_FINAL, first, 1, rbase, NULL);
        else if (nch == 2)
            encode_pcm_block(&E, &E.ss[0], base, base + 1, 2, n, done, F_INITIAL | F_FINAL, first, 2, rbase, rbase ? rbase + 1 : NULL);
        else { /* 5.1: FL/FR stereo, FC mono, LFE mono, BL/BR stereo */
            encode_pcm_block(&E, &E.ss[0], base, base + 1, 6, n, done, F_INITIAL, first, 6, rbase, rbase ? rbase + 1 : NULL);
            encode_pcm_block(&E, &E.ss[1], base + 2, NULL, 6, n, done, 0, 0, 6, rbase ? rbase + 2 : NULL, NULL);
            encode_pcm_block(&E, &E.ss[2], base + 3, NULL, 6, n, done, 0, 0, 6, rbase ? rbase + 3 : NULL, NULL);
            encode_pcm_block(&E, &E.ss[3], base + 4, base + 5, 6, n, done, F_FINAL, 0, 6, rbase ? rbase + 4 : NULL, rbase ? rbase + 5 : NULL);
        }
        done += n;
        first = 0;
        if (E.out.overflow) break;
    }
    if (!E.out.overflow && (cfg->extras & WVENC_X_MD5_TRAILER)) {
        /* metadata-only block: block_samples == 0 (WavPackUtils.cs:219) */
        obuf *o = &E.out;
        size_t hdr_at = o->len;
        u32 flags = (u32)((cfg->kind == WVENC_DSD ? 0 : (cfg->bits + 7) / 8 - 1)) | F_INITIAL | F_FINAL | (nch == 1 ? F_MONO : 0) |
                    ((u32)srate_index(cfg->sample_rate) << 23) | (cfg->kind == WVENC_DSD ? F_DSD : 0);
        header_put(o, 0, E.version, E.total_field, nsamples, 0, flags, 0xffffffffu);
        uint8_t md5[16];
        for (int i = 0; i < 16; i++) md5[i] = (uint8_t)(i * 17 + 3);
        put_meta(o, 0x26, md5, 16);
        put_meta(o, 0x22, (const uint8_t *)"LISTtrailer!", 12);
        if (!o->overflow) { u32 cks = (u32)(o->len - hdr_at - 8); memcpy(o->p + hdr_at + 4, &cks, 4); }
    }
    free(E.bits);
    free(E.wvx);
    return E.out.overflow ? 0 : E.out.len;
}

/* ------------------------------------------------------------------ */
/* multi-threaded corpus                                                */
/* ------------------------------------------------------------------ */
typedef struct {
    const wvenc_config *cfg;
    i64 nsamples, nfiles;
    u64 base_seed;
    uint8_t *out; size_t cap;
    u64 *offsets, *sizes;
    atomic_llong next_file;
    atomic_ullong bump;
    atomic_int failed;
} corpus_job;

static void *corpus_worker(void *arg)
{
    corpus_job *J = (corpus_job *)arg;
    size_t bound = wvenc_bound(J->cfg, J->nsamples);
    uint8_t *tmp = (uint8_t *)malloc(bound);
    i32 *pcm = (i32 *)malloc(sizeof(i32) * (size_t)(J->nsamples + 1) * J->cfg->channels);
    for (;;) {
        long long i = atomic_fetch_add(&J->next_file, 1);
        if (i >= J->nfiles || atomic_load(&J->failed)) break;
        wvenc_synth(J->cfg, J->base_seed + (u64)i, J->nsamples, pcm);
        size_t len = wvenc_encode(J->cfg, pcm, J->nsamples, tmp, bound, NULL);
        if (!len) { atomic_store(&J->failed, 1); break; }
        size_t alen = (len + 63) & ~(size_t)63;
        unsigned long long at = atomic_fetch_add(&J->bump, alen);
        if (at + alen > J->cap) { atomic_store(&J->failed, 1); break; }
        memcpy(J->out + at, tmp, len);
        memset(J->out + at + len, 0, alen - len);
        J->offsets[i] = at;
        J->sizes[i] = len;
    }
    free(tmp);
    free(pcm);
    return NULL;
}

size_t wvenc_build_corpus(const wvenc_config *cfg, int64_t nsamples_per_file, int64_t nfiles, uint64_t base_seed, int threads,
                          uint8_t *out, size_t cap, uint64_t *offsets, uint64_t *sizes)
{
    init_tables();
    if (threads <= 0) { long n = sysconf(_SC_NPROCESSORS_ONLN); threads = n > 0 ? (int)n : 1; }
    if (threads > 256) threads = 256;
    if (threads > nfiles) threads = (int)(nfiles > 0 ? nfiles : 1);
    corpus_job J;
    memset(&J, 0, sizeof(J));
    J.cfg = cfg; J.nsamples = nsamples_per_file; J.nfiles = nfiles; J.base_seed = base_seed;
    J.out = out; J.cap = cap; J.offsets = offsets; J.sizes = sizes;
    atomic_init(&J.next_file, 0); atomic_init(&J.bump, 0); atomic_init(&J.failed, 0);
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, corpus_worker, &J);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    if (atomic_load(&J.failed)) return 0;
    return (size_t)atomic_load(&J.bump);
}
