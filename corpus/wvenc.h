/*
 * corpus/wvenc.h -- synthetic WavPack corpus generator (bench/test infrastructure).
 *
 * No datasets are available offline, so the bench and the tests make their own
 * .wv files: an integer-only signal generator plus a small WavPack encoder that
 * emits the stream features the reference decoder understands.
 * This is NOT part of the product decode path and NOT part of the oracle.  The
 * header CRC of lossless blocks is computed from the SOURCE samples, which makes
 * it an independent check of any decoder.
 */
#ifndef WVENC_H
#define WVENC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    WVENC_PCM = 0,      /* lossless integer PCM */
    WVENC_HYBRID = 1,   /* hybrid lossy (HYBRID_FLAG|HYBRID_BITRATE) */
    WVENC_FLOAT = 2,    /* FLOAT_DATA flag + ID_FLOAT_INFO + dummy WVX (reference decodes mantissa ints) */
    WVENC_DSD = 3       /* DSD_FLAG blocks, dsd_mode 0/1/3 */
};

enum { /* extras bitmask */
    WVENC_X_RIFF_HEADER = 1,    /* ID_RIFF_HEADER in first block */
    WVENC_X_CONFIG = 2,         /* ID_CONFIG_BLOCK in first block */
    WVENC_X_NEW_CONFIG = 4,     /* ID_NEW_CONFIG_BLOCK (WavPack 5) */
    WVENC_X_BLOCK_CHECKSUM = 8, /* ID_BLOCK_CHECKSUM at the end of every block */
    WVENC_X_MD5_TRAILER = 16,   /* final block_samples==0 block carrying ID_MD5_CHECKSUM + ID_RIFF_TRAILER */
    WVENC_X_SAMPLE_RATE = 32,   /* SRATE index 15 + ID_SAMPLE_RATE */
    WVENC_X_DUMMY = 64,         /* an ID_DUMMY sub-block and an odd-sized optional sub-block */
    WVENC_X_ALL_HISTORY = 128   /* store decorr history for every pass (exercises quirk C-1) */
};

typedef struct {
    int32_t kind;            /* WVENC_* */
    int32_t sample_rate;
    int32_t bits;            /* source bits per sample: 8,16,24,32 (PCM); 32 for FLOAT; 8 for DSD */
    int32_t channels;        /* 1, 2, or 6 (5.1: stereo+mono+mono+stereo blocks per segment) */
    int32_t block_samples;   /* samples per block (DSD: byte-times per block) */
    int32_t nterms;          /* 0..16 */
    int8_t terms[16];        /* file (encoder) order; stereo-only terms are replaced for mono blocks */
    int8_t deltas[16];
    int32_t joint_stereo;    /* 1: JOINT_STEREO on stereo blocks */
    int32_t false_stereo;    /* 1: emit FALSE_STEREO blocks when L==R over a block */
    int32_t shift;           /* header SHIFT field (low zero bits removed), 0..  */
    /* INT32_DATA handling for bits==32 */
    int32_t int32_sent_bits; /* >0: INT32_DATA with that many low bits in WVX */
    int32_t int32_wvx;       /* 1: write the WVX stream (lossless); 0: omit it (lossy) */
    int32_t int32_new_wvx;   /* 1: ID_WVX_NEW_BITSTREAM with int32_max_width field */
    int32_t int32_max_width; /* value for the 5-bit field when int32_new_wvx */
    int32_t int32_zeros, int32_ones, int32_dups; /* redundancy fields (mutually exclusive) */
    /* hybrid */
    int32_t hybrid_bitrate;  /* bits per sample * 256 (e.g. 4*256) */
    int32_t hybrid_balance;  /* 1: HYBRID_BALANCE */
    /* float */
    int32_t float_flags, float_shift, float_max_exp, float_norm_exp;
    int32_t float_new_wvx;
    /* dsd */
    int32_t dsd_mode;        /* 0,1,3 */
    int32_t dsd_rate_shift;  /* first byte of ID_DSD_BLOCK (dsd_multiplier = 1<<this) */
    int32_t dsd_history_bits;/* mode 1: 0..5 */
    int32_t dsd_raw_probs;   /* mode 1: 1 -> max_probability=0xFF raw table path */
    int32_t dsd_rate_i;      /* mode 3: ptable rate */
    int32_t extras;          /* WVENC_X_* */
    int32_t version;         /* stream version, 0 -> 0x410 */
    int32_t unknown_length;  /* 1: total_samples = 0xFFFFFFFF */
} wvenc_config;

void wvenc_default_config(wvenc_config *cfg); /* 16-bit stereo 44.1k, terms {18,18,2,3,-2} delta 2, joint stereo */

/* Integer-only synthetic signal (PCG32; seed = 0x5EED0000 + file_id by convention).
 * out: interleaved int32, nsamples*channels entries, range of `bits`.
 * For DSD: out holds one byte value (0..255) per channel-sample (a 1-bit
 * sigma-delta modulation of the same signal, MSB first). */
void wvenc_synth(const wvenc_config *cfg, uint64_t seed, int64_t nsamples, int32_t *out);

/* Encode interleaved samples into a complete .wv byte stream.  Returns bytes
 * written, or 0 if cap is too small / config invalid.  recon (optional,
 * nsamples*channels) receives what a correct decoder must output as
 * right-justified int32 (for lossless: the source; hybrid/float/lossy-int32:
 * the simulated decoder output incl. fixup). */
size_t wvenc_encode(const wvenc_config *cfg, const int32_t *samples, int64_t nsamples, uint8_t *out, size_t cap, int32_t *recon);

/* Worst-case output size for wvenc_encode. */
size_t wvenc_bound(const wvenc_config *cfg, int64_t nsamples);

/* Multi-threaded corpus build: nfiles files, file i synthesised with seed
 * base_seed+i and encoded; file i is written at out + offsets[i] (offsets and
 * sizes are outputs; files are packed back to back with 64-byte alignment).
 * Returns total bytes used, 0 on overflow. threads<=0 -> hardware concurrency. */
size_t wvenc_build_corpus(const wvenc_config *cfg, int64_t nsamples_per_file, int64_t nfiles, uint64_t base_seed, int threads,
                          uint8_t *out, size_t cap, uint64_t *offsets, uint64_t *sizes);

#ifdef __cplusplus
}
#endif
#endif
