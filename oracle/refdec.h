/*
 * oracle/refdec.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C restatement of the Quake4/WavPackDecoder (C#) decode path, used as the
 * parity oracle and as the timed CPU baseline ("port").  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path (wavpackdecoder_b200/) never does.
 *
 * Pinning status: the reference ships no golden vectors and cannot be executed
 * in this environment (no .NET/mono), so parity with the C# original is
 * UNPINNED by reference-owned vectors.  The oracle is instead cross-pinned
 * against an independent WavPack implementation (FFmpeg libavcodec 62.11 native
 * encoder + decoder, see tests/golden/README.md) and against source-sample
 * CRCs written by the in-repo encoder.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).
 */
#ifndef REFDEC_H
#define REFDEC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rd_context rd_context;

/* Defines.cs:18,26 */
#define RD_SAMPLE_BUFFER_SIZE 4096
#define RD_OPEN_2CH_MAX 0x8

/* WavPackUtils.cs:36  WavpackOpenFileInput over an in-memory stream.
 * Never returns NULL; inspect rd_get_error_message(). */
rd_context *rd_open(const uint8_t *file, size_t len, uint32_t flags);
void rd_close(rd_context *c);

/* WavPackUtils.cs:200.  Returns samples unpacked, or -2 if the C# code would
 * have thrown (IndexOutOfRange / EndOfStream) out of the call. */
long rd_unpack_samples(rd_context *c, int32_t *buffer, long buffer_len, long samples);

/* WavPackUtils.cs:504-594  SetSample / SetTime -> seek().  1 / 0 like the C# bool; -2 where the C# code would throw out
 * of the call; -3 where it would loop forever (WavpackUnpackSamples returning 0 inside the skip loop). */
int rd_set_sample(rd_context *c, long sample);
int rd_set_time(rd_context *c, long milliseconds);

/* WavPackUtils.cs:288 (static, context-free). returns 1/0 like the C# bool. */
int rd_format_samples(const int32_t *src, long samcnt, int bps, uint8_t *pcm, long pcm_len, int offset, int dsd);

/* getters, WavPackUtils.cs:133-499 */
long rd_get_num_samples(rd_context *c, int native);
long rd_get_sample_index(rd_context *c);
long rd_get_num_errors(rd_context *c);
int rd_lossy(rd_context *c);
long rd_get_sample_rate(rd_context *c);
int rd_get_num_channels(rd_context *c);
int rd_get_bits_per_sample(rd_context *c);
int rd_get_bytes_per_sample(rd_context *c);
int rd_get_reduced_channels(rd_context *c);
int rd_get_file_format(rd_context *c);
const char *rd_get_file_extension(rd_context *c);
const char *rd_get_error_message(rd_context *c); /* NULL when none */
const uint8_t *rd_get_header(rd_context *c, long *len);
const uint8_t *rd_get_trailer(rd_context *c, long *len);
int rd_get_is_five(rd_context *c);
int rd_get_version(rd_context *c);
int rd_get_is_float(rd_context *c);
int rd_get_mode(rd_context *c);
/* writes "" when the C# returns null */
void rd_get_compression_level(rd_context *c, char *out, size_t cap);

/* Oracle-only helpers (not in the reference API). */
/* CRC accumulated for the block most recently finished / in progress. */
int32_t rd_dbg_block_crc(rd_context *c);
int rd_dbg_mute_error(rd_context *c);
uint32_t rd_dbg_block_flags(rd_context *c);
/* Decode the block the context is positioned on (after rd_open at that block)
 * with unpack_samples/unpack_dsd_samples directly, ignoring INITIAL_BLOCK, in
 * `chunk`-sample calls.  Used to pin non-initial multichannel blocks.
 * Returns samples decoded (block_samples) or -2 on exception. */
long rd_dbg_unpack_current_block(rd_context *c, int32_t *buffer, long buffer_len, long chunk);
/* 1 if the current block failed check_crc_error (UnpackUtils.cs:1414) */
int rd_dbg_check_crc_error(rd_context *c);

/* Whole-file convenience used by tests and the CPU baseline: decode like
 * WvDemo.cs:110-135 (chunked Unpack -> Format) into pcm.  Returns total
 * samples unpacked or <0.  md5 is not computed here. */
long rd_decode_file_pcm(const uint8_t *file, size_t len, uint32_t open_flags, long chunk_samples,
                        uint8_t *pcm, size_t pcm_cap, size_t *pcm_len, long *crc_errors);

/* WvDemo.Main restated (WvDemo.cs:15-174): writes the bytes the demo writes to its output file into out and its
 * process exit code into *exit_code.  Returns the byte count, -1 if cap is too small. */
long rd_wvdemo(const uint8_t *file, size_t len, uint8_t *out, size_t cap, int *exit_code);

#ifdef __cplusplus
}
#endif
#endif
