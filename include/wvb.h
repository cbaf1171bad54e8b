/*
 * wvb.h -- C ABI of the B200-native batch WavPack block decoder (libwvb.so).
 *
 * This is the drop-in boundary for the reference's decode path.  Each entry point
 * names the reference interface it replaces (paths relative to the reference tree
 * Quake4/WavPackDecoder).  Plain C: pointers, sizes, integer status codes; no
 * exceptions, no torch types, no callbacks into the host language.  The host-side
 * mirrors of the reference API (C#: csharp/WavPackUtils.cs via P/Invoke, Python:
 * wavpackdecoder_b200.wavpack_utils via ctypes) are thin layers over these calls.
 *
 * Division of labour (reference: WavPackUtils.cs:200-282 drives everything serially):
 *   host   wvb_index*        header hop + sub-block (TLV) walk + config ids  -> block table
 *   device wvb_batch_decode  per block: metadata parse, get_words, decorrelation,
 *                            joint stereo, CRC, mute check, fixup/float/int32,
 *                            FALSE_STEREO expansion, int32 or packed-PCM stores;
 *                            DSD modes 0/1/3
 * There is no CPU decode fallback: without a CUDA device wvb_batch_* return
 * WVB_E_NO_DEVICE.
 */
#ifndef WVB_H
#define WVB_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WVB_ABI_VERSION 3 /* 3: wvb_block_desc.checksum_off (was reserved), WVB_BF_BLOCK_CHECKSUM / WVB_RF_BLOCK_CHECKSUM, wvb_batch_decode_files, wvb_batch_dsd_to_dsf, wvb_block_checksum_ok */

/* status codes */
enum {
    WVB_OK = 0,
    WVB_E_ARG = -1,        /* bad argument */
    WVB_E_NO_DEVICE = -2,  /* no CUDA device / driver */
    WVB_E_CUDA = -3,       /* CUDA runtime error, see wvb_last_error() */
    WVB_E_CAPACITY = -4,   /* output table too small */
    WVB_E_FORMAT = -5      /* not a WavPack stream the reference would open */
};

/* open flags: Defines.cs:26 (OPEN_2CH_MAX) plus extensions above bit 16 */
#define WVB_OPEN_2CH_MAX 0x8u
#define WVB_OPEN_ALL_CHANNELS 0x10000u /* extension: also index the non-INITIAL blocks of a segment */

/* output formats of wvb_batch_decode */
enum {
    WVB_OUT_INT32 = 0, /* right-justified int32, what WavpackUnpackSamples returns (WavPackUtils.cs:190-198) */
    WVB_OUT_PCM = 1    /* little-endian packed PCM of bytes_per_sample bytes, what WavpackFormatSamples
                          produces with dsd=false (WavPackUtils.cs:288-341; 8-bit gets +128) */
};
#define WVB_OUT_DSD_RAW 2 /* like WVB_OUT_PCM with bps=1 but bytes copied raw (WavpackFormatSamples dsd:true) */

/* memory-space flags for wvb_batch_decode */
#define WVB_IN_DEVICE 1u   /* `in` is a device pointer (already resident in HBM) */
#define WVB_OUT_DEVICE 2u  /* `out` is a device pointer */
#define WVB_RESULTS_DEVICE 4u /* `results` is a device pointer */
#define WVB_NO_SYNC 8u     /* return after enqueueing; call wvb_batch_wait() */

/* sub-block slots recorded by the index pass (Defines.cs:57-69 ids) */
enum {
    WVB_SUB_TERMS = 0,   /* ID_DECORR_TERMS   0x02 */
    WVB_SUB_WEIGHTS = 1, /* ID_DECORR_WEIGHTS 0x03 */
    WVB_SUB_SAMPLES = 2, /* ID_DECORR_SAMPLES 0x04 */
    WVB_SUB_ENTROPY = 3, /* ID_ENTROPY_VARS   0x05 */
    WVB_SUB_HYBRID = 4,  /* ID_HYBRID_PROFILE 0x06 */
    WVB_SUB_WV = 5,      /* ID_WV_BITSTREAM   0x0a */
    WVB_SUB_WVX = 6,     /* ID_WVX_BITSTREAM  0x0c / ID_WVX_NEW_BITSTREAM 0x2c */
    WVB_SUB_DSD = 7,     /* ID_DSD_BLOCK      0x0e */
    WVB_SUB_COUNT = 8
};

/* block-descriptor flags (wvb_block_desc.bflags) */
#define WVB_BF_WVX_NEW 1u        /* WVX sub-block used the new id (two/one 5-bit fields first) */
#define WVB_BF_HAS_INT32_INFO 2u /* int32_info valid (possibly inherited from an earlier block, quirk C-8) */
#define WVB_BF_HAS_FLOAT_INFO 4u
#define WVB_BF_WVX_PRESENT 8u    /* wps.wvxbits != null at this block (possibly inherited) */
#define WVB_BF_MUTE_ALL 16u      /* reference reaches this block without unpack_init (after a gap): output zeros, count a CRC error */
#define WVB_BF_STALE_STATE 32u   /* block depends on decoder state left by an earlier block; result flagged inexact */
#define WVB_BF_DSD_PADDED 64u    /* DSD payload array includes the pad byte (data.Length quirk C-10) */
#define WVB_BF_BLOCK_CHECKSUM 128u /* the block ends with a well-formed ID_BLOCK_CHECKSUM sub-block at checksum_off */

/*
 * One decodable block, as produced by wvb_index (file-relative offsets) and consumed
 * by wvb_batch_decode (slab-absolute offsets, see wvb_rebase).  160 bytes.
 * Replaces: WavpackHeader (WavpackHeader.cs:15-22) + the WavpackMetadata walk of
 * unpack_init (UnpackUtils.cs:24-68, MetadataUtils.cs:15-109).
 */
typedef struct wvb_block_desc {
    uint64_t in_offset;     /* byte offset of the 32-byte 'wvpk' header */
    uint64_t out_offset;    /* wvb_index: first output sample index (complete samples) within the file;
                               wvb_rebase turns it into a byte offset in the output slab */
    uint32_t in_bytes;      /* ckSize + 8 */
    uint32_t block_samples;
    uint32_t flags;         /* header flags (Defines.cs:28-44) */
    int32_t crc;            /* header crc */
    int64_t block_index;    /* header block_index (40 bit) */
    uint32_t sub_off[WVB_SUB_COUNT]; /* payload offset relative to in_offset, 0 = absent */
    uint32_t sub_len[WVB_SUB_COUNT]; /* byte_length (ODD_SIZE already applied) */
    uint8_t int32_info[4];  /* sent_bits, zeros, ones, dups (UnpackUtils.cs:376-379) */
    uint8_t float_info[4];  /* flags, shift, max_exp, norm_exp (FloatUtils.cs:23-26) */
    uint32_t bflags;        /* WVB_BF_* */
    uint16_t version;
    uint8_t out_channels;   /* channels this block writes (FALSE_STEREO -> 2) */
    uint8_t out_stride;     /* channels per output frame (== out_channels unless WVB_OPEN_ALL_CHANNELS) */
    uint8_t out_ch_offset;  /* first channel slot inside the frame */
    uint8_t out_bps;        /* bytes per sample for WVB_OUT_PCM (config.bytes_per_sample of the file) */
    uint16_t smem_words;    /* per-thread shared-memory words the PCM kernel needs for this block's decorrelation state */
    uint32_t chunk_first;   /* samples from block start to the caller's next chunk boundary (mute/short-weight semantics) */
    uint32_t chunk_samples; /* caller chunk size in samples (Defines.cs:18 SAMPLE_BUFFER_SIZE = 4096 in WvDemo) */
    uint32_t file_id;       /* caller tag */
    uint32_t gap_before;    /* samples of zero fill before this block (WavPackUtils.cs:227-251) */
    uint32_t terms_sig;     /* hash of (terms, deltas): the planner groups equal signatures into the same warps */
    uint32_t skip_samples;  /* wvb_index_seek: leading samples of the block that seek() decodes and discards in pieces of ... */
    uint32_t skip_chunk;    /* ... this many samples (SAMPLE_BUFFER_SIZE / channels, WavPackUtils.cs:573-578); both 0 otherwise */
    uint32_t avg_block_size; /* wphdr.average_block_size after this block's header was read (WavPackUtils.cs:647-650): the
                                reference's seek() extrapolates file positions from it */
    uint32_t checksum_off;  /* WVB_BF_BLOCK_CHECKSUM: offset of the ID_BLOCK_CHECKSUM sub-block (its id byte) from in_offset */
} wvb_block_desc;

/* per-block result.  Replaces wps.crc / mute_error / check_crc_error (UnpackUtils.cs:1414-1421). 16 bytes */
#define WVB_RF_CRC_ERROR 1u     /* check_crc_error() would be true -> wpc.crc_errors++ (WavPackUtils.cs:273-275) */
#define WVB_RF_MUTED 2u         /* mute_error latched (UnpackUtils.cs:649-664 / DsdUtils.cs:99-117) */
#define WVB_RF_CRCX_ERROR 4u    /* extended (WVX) crc mismatch */
#define WVB_RF_INEXACT 8u       /* corrupt-stream corner the device path does not reproduce bit-exactly */
#define WVB_RF_BAD_BLOCK 16u    /* device-side metadata validation failed (DSD tables) */
#define WVB_RF_BLOCK_CHECKSUM 32u /* the WavPack 5 block checksum does not match the block's bytes.  An extension: the reference
                                     only notes the sub-block (MetadataUtils.cs:183), so output and crc_errors are unaffected */
typedef struct wvb_block_result {
    int32_t crc;           /* crc accumulated by the decoder */
    uint32_t rflags;       /* WVB_RF_* */
    uint32_t mute_from;    /* first muted sample of the block when WVB_RF_MUTED */
    int32_t crc_x;
} wvb_block_result;

/*
 * File-level information.  Replaces the WavpackContext/WavpackConfig state built by
 * WavpackOpenFileInput (WavPackUtils.cs:36-120) and read by the getters
 * (WavPackUtils.cs:133-499).
 */
typedef struct wvb_file_info {
    int32_t status;            /* WVB_OK or WVB_E_FORMAT */
    char error_message[64];    /* the reference's error_message text, "" if none */
    int64_t total_samples;     /* wpc.total_samples, -1 unknown */
    int64_t sample_rate;       /* config.sample_rate (before the DSD multiplier) */
    int64_t config_flags;      /* config.flags */
    int64_t channel_mask;
    int32_t num_channels;      /* config.num_channels */
    int32_t reduced_channels;  /* wpc.reduced_channels (0 if unset) */
    int32_t bits_per_sample;   /* config.bits_per_sample */
    int32_t bytes_per_sample;  /* config.bytes_per_sample */
    int32_t float_norm_exp;
    int32_t xmode;
    int32_t version;           /* first block's stream version */
    int32_t five;              /* wpc.five */
    int32_t file_format;       /* eFileFormat */
    int32_t lossy_blocks;      /* wpc.lossy_blocks after indexing every block */
    uint32_t dsd_multiplier;
    uint32_t first_flags;      /* flags of the first audio block */
    int64_t header_off, header_len;   /* stored RIFF/alt header bytes (file offsets), len -1 if none */
    int64_t trailer_off, trailer_len;
    char file_extension[16];   /* "" -> "wav" */
    int64_t num_blocks;        /* descriptors written */
    int64_t indexed_samples;   /* complete samples covered by the descriptors incl. gaps */
    int32_t stopped_early;     /* 1: the reference's sequential reader would stop (bad metadata / lost sync) before EOF */
    int32_t reserved;
} wvb_file_info;

/* ---- library ---- */
int wvb_abi_version(void);
/* The struct layouts this build was compiled with, as text: "struct:size;field:offset:size;...|struct:..." for
 * wvb_block_desc, wvb_block_result, wvb_file_info and wvb_seek_state.  Host-language bindings that mirror the structs by
 * hand (ctypes, C# StructLayout) compare themselves with it at start-up or in their tests. */
const char *wvb_abi_layout(void);
const char *wvb_last_error(void); /* thread-local text for the last non-OK status */
int wvb_device_count(void);       /* 0 without a driver/device */

/* ---- host index pass: replaces read_next_header (WavPackUtils.cs:600-671), unpack_init's
 * metadata walk (UnpackUtils.cs:24-68) and the config part of WavpackOpenFileInput ---- */
/* Index one in-memory .wv file.  chunk_samples = the size the caller would pass to
 * WavpackUnpackSamples (4096 in WvDemo); it only matters for corrupt streams.
 * blocks may be NULL with cap 0 to count. */
int wvb_index(const uint8_t *file, size_t len, uint32_t open_flags, uint32_t chunk_samples, wvb_file_info *info,
              wvb_block_desc *blocks, size_t cap, size_t *nblocks);

/* Seek (replaces seek(), WavPackUtils.cs:521-594, reached through SetSample / SetTime :502-512).
 *
 * The reference's reader state a seek starts from: the header it read last and where the file pointer stands.  From a
 * descriptor of the block the caller read last (or the first block, if nothing was read yet): hdr_pos = in_offset,
 * ck_size = in_bytes - 8, block_index / block_samples / avg_block_size as in the descriptor, file_pos = in_offset + in_bytes. */
typedef struct wvb_seek_state {
    int64_t hdr_pos;        /* wphdr.stream_position */
    int64_t block_index;
    int64_t avg_block_size; /* wphdr.average_block_size */
    int64_t file_pos;       /* infile.BaseStream.Position */
    uint32_t block_samples;
    uint32_t ck_size;
} wvb_seek_state;

/* Index the file as the reference decodes it AFTER seek() to complete sample `target` followed by WavpackUnpackSamples
 * calls of `chunk_samples`.  The reference probes for a block whose header range contains the target (at most 25 header
 * reads steered by the average block size), restarts its decoder there with a fresh stream state, and decodes and
 * discards that block's samples before the target in pieces of `skip_chunk` samples (SAMPLE_BUFFER_SIZE / reduced
 * channels; 0 selects that); the caller's call grid starts at the target.
 *   from != NULL: the probe sequence is replayed on the headers exactly (including its quirks: a target that is the first
 *                 sample of a block is usually reached by decoding the whole previous block, whose CRC verdict then counts;
 *                 when the probes run out the last header read is taken, wherever it is);
 *   from == NULL: the block is found by a header hop from the start of the file (what a new caller wants).
 * Descriptors come back for the block decoding restarts at and up to max_blocks - 1 following ones (0: to the end of the
 * file), out_offset counted from that block's first sample (*window_first_sample in the file).  *landed_sample is the
 * sample index the reader stands at afterwards (the target, except after a failed probe sequence): the caller's data begins
 * *landed_sample - *window_first_sample samples into the decoded window.
 * info describes the FILE (as wvb_index), except num_blocks / indexed_samples / stopped_early, which describe the window.
 * Returns WVB_OK with *nblocks == 0 where the reference's seek() returns false (target past the end, unknown length, a
 * probe before the start of the file) or -- from == NULL -- when no block contains the target. */
int wvb_index_seek(const uint8_t *file, size_t len, uint32_t open_flags, const wvb_seek_state *from, int64_t target, uint32_t skip_chunk,
                   uint32_t chunk_samples, size_t max_blocks, wvb_file_info *info, wvb_block_desc *blocks, size_t cap, size_t *nblocks,
                   int64_t *window_first_sample, int64_t *landed_sample);

/* Sizes are the stream's word: like the reference, the index pads a gap in block_index (or a total_samples the blocks never
 * reach) with zeros, so a damaged or hostile header can ask for billions of output samples.  Callers that decode untrusted
 * files should bound wvb_file_info.indexed_samples / the out_bytes of wvb_index_many before allocating. */
/* Index many files with `threads` host threads (<=0: all cores).  File i occupies
 * slab[offsets[i] .. offsets[i]+sizes[i]); its descriptors are written to
 * blocks[first[i] .. first[i]+count[i]) (first/count are outputs) already rebased
 * for a file-major output slab in out_format.  out_bytes receives the output slab size.
 * blocks == NULL (cap 0) only counts.  With a table, the files are walked once; if cap turns out too small the call
 * returns WVB_E_CAPACITY with *nblocks set to the size needed (callers that decode similar batches repeatedly pass the
 * previous count as cap and skip the counting call). */
int wvb_index_many(const uint8_t *slab, const uint64_t *offsets, const uint64_t *sizes, size_t nfiles, uint32_t open_flags,
                   uint32_t chunk_samples, int out_format, int threads, wvb_file_info *infos, wvb_block_desc *blocks, size_t cap,
                   uint64_t *first, uint64_t *count, uint64_t *file_out_offset, size_t *nblocks, uint64_t *out_bytes);

/* Turn file-relative descriptors into slab-absolute ones: in_offset += in_base;
 * out_offset = out_base + out_offset(samples) * out_stride * unit_bytes.
 * Alignment of out_base: int32 output needs a multiple of 4; packed PCM accepts any offset (block outputs are
 * assembled into aligned words with bytewise first/last words), 16-bit stereo decodes fastest at multiples of 4. */
void wvb_rebase(wvb_block_desc *blocks, size_t n, uint64_t in_base, uint64_t out_base, int out_format, uint32_t file_id);

/* bytes one complete sample occupies in the output for this descriptor */
uint32_t wvb_frame_bytes(const wvb_block_desc *b, int out_format);

/* ---- device batch decode: replaces the body of WavpackUnpackSamples (WavPackUtils.cs:200-282):
 * unpack_samples / unpack_dsd_samples / check_crc_error, and WavpackFormatSamples ---- */
typedef struct wvb_batch wvb_batch;
int wvb_batch_create(int device, wvb_batch **out);
void wvb_batch_destroy(wvb_batch *b);

/* Decode nblocks blocks.  The slab passed as `in` must have at least 16 readable bytes after the last block (the
 * kernels fetch whole aligned words one word ahead); host slabs are copied into padded device memory by the library,
 * device slabs (WVB_IN_DEVICE) must be allocated with that slack.  `in` holds the compressed slab (descs' in_offset index it), `out`
 * receives int32 or packed PCM at descs' out_offset.  Host pointers are copied through the
 * batch's device buffers (pinned host memory makes the copies asynchronous); device pointers
 * (WVB_*_DEVICE) are used in place.  descs is always a host pointer.  results may be NULL.
 * Large host-input batches (>= ~3 GB of input + output, table in slab order) are decoded in segments so that the
 * upload, the kernels and -- for host output -- the download overlap; with WVB_OUT_DEVICE the PCM stays on the device
 * (the verify flow, see wvb_batch_md5). */
int wvb_batch_decode(wvb_batch *b, const uint8_t *in, size_t in_bytes, const wvb_block_desc *descs, size_t nblocks, void *out,
                     size_t out_bytes, int out_format, uint32_t mem_flags, wvb_block_result *results);
int wvb_batch_wait(wvb_batch *b);

/* Upload a block table once (descriptors + launch plan) so that repeated decodes of the same
 * table skip the planning and the descriptor copy: afterwards call wvb_batch_decode with
 * descs == NULL and the same nblocks.  This is the "open" half of the reference's open/decode
 * split (the table is what WavpackOpenFileInput + the header reads would have produced). */
int wvb_batch_prepare(wvb_batch *b, const wvb_block_desc *descs, size_t nblocks, int out_format);

/* Device time of the last wvb_batch_decode, from CUDA events on the batch's stream:
 * kernel_ms = decode kernels only, h2d_ms/d2h_ms = the copies, launches = kernels launched. */
int wvb_batch_timing(wvb_batch *b, float *kernel_ms, float *h2d_ms, float *d2h_ms, int *launches);
/* the batch's cudaStream_t, for callers that want to order their own work after it */
void *wvb_batch_stream(wvb_batch *b);

/* ---- integrity (beyond the reference, which ignores ID_MD5_CHECKSUM, MetadataUtils.cs:187-191) ----
 * MD5 (RFC 1321) of n byte ranges of a decoded output slab, computed on the device: digests[16*i..] = MD5 of
 * slab[offsets[i] .. offsets[i]+lengths[i]).  device_out: the slab a WVB_OUT_DEVICE decode wrote (out_bytes = its
 * size), or NULL for the batch's own device copy of the last host-buffer decode.  offsets/lengths/digests are host
 * pointers.  With one range per file (its PCM) the result compares directly with the file's stored MD5, and a
 * verify-only pass never moves PCM over PCIe. */
int wvb_batch_md5(wvb_batch *b, const void *device_out, size_t out_bytes, const uint64_t *offsets, const uint64_t *lengths, size_t n,
                  uint8_t *digests);
/* The MD5 a .wv file stores for its source audio (ID_MD5_CHECKSUM, Defines.cs:77), searched in every block's
 * metadata.  Returns 1 and fills md5 if present, 0 if not. */
int wvb_stored_md5(const uint8_t *file, size_t len, uint8_t md5[16]);

/* The WavPack 5 block checksum of one block (ID_BLOCK_CHECKSUM, Defines.cs:83), on the host: 1 = present and matching,
 * 0 = present and wrong, -1 = the block has none (or is malformed).  Definition (WavPack 5 libwavpack,
 * WavpackVerifySingleBlock): csum = 0xffffffff; csum = csum * 3 + w over the 16-bit little-endian words of the block from
 * its 'wvpk' up to the checksum sub-block; stored as 4 bytes, or as the 2 bytes of csum ^ (csum >> 16).  wvb_batch_decode
 * checks the same thing on the device for every block that has one (a warp per block, after the decode kernels) and reports
 * WVB_RF_BLOCK_CHECKSUM. */
int wvb_block_checksum_ok(const uint8_t *block, size_t len);

/* Index AND decode a slab of files in one call (host input, host or device output).  Same results as wvb_index_many followed
 * by wvb_batch_decode -- the same table layout (first / count / file_out_offset, descriptors rebased for a file-major
 * output slab with every file 16-byte aligned), the same output bytes and block results -- but the index pass overlaps the
 * upload and decode: files are indexed in slab order by `threads` host threads while the segments whose files are already
 * indexed are uploaded, decoded and downloaded.  offsets[] must be ascending and the files disjoint.  `blocks` (cap
 * entries) receives the table, `out` (out_cap bytes) the samples.  If cap or out_cap is too small the contents of `out` are
 * unspecified, *nblocks / *out_bytes hold the sizes needed and the call returns WVB_E_CAPACITY (callers that decode similar
 * batches repeatedly keep the previous sizes).  mem_flags: WVB_OUT_DEVICE and WVB_NO_SYNC only.
 * Replaces the open / unpack loop of WvDemo.cs:41-141 run over many files. */
int wvb_batch_decode_files(wvb_batch *b, const uint8_t *slab, size_t slab_bytes, const uint64_t *offsets, const uint64_t *sizes, size_t nfiles,
                           uint32_t open_flags, uint32_t chunk_samples, int out_format, int threads, wvb_file_info *infos,
                           wvb_block_desc *blocks, size_t cap, uint64_t *first, uint64_t *count, uint64_t *file_out_offset, size_t *nblocks,
                           uint64_t *out_bytes, void *out, size_t out_cap, uint32_t mem_flags, wvb_block_result *results);

/* Container writer, DSD (beyond the reference, whose demo writes RIFF only, WvDemo.cs:78-105): re-lay decoded DSD bytes
 * (WVB_OUT_DSD_RAW: one byte per channel per byte-time, interleaved, oldest bit in the MSB -- the DSDIFF layout) into the
 * Sony DSF layout (4096-byte blocks per channel, oldest bit in the LSB, last block zero padded) on the device.  File i:
 * frames[i] byte-times of channels[i] channels at device_src + src_off[i]  ->  ceil(frames/4096) * 4096 * channels bytes
 * at device_dst + dst_off[i] (4-byte aligned).  Both slabs are device memory. */
int wvb_batch_dsd_to_dsf(wvb_batch *b, const void *device_src, size_t src_bytes, void *device_dst, size_t dst_bytes, const uint64_t *src_off,
                         const uint64_t *dst_off, const uint64_t *frames, const uint32_t *channels, size_t nfiles);

/* pinned host memory helpers for hosts without their own allocator (C# shim) */
void *wvb_host_alloc(size_t bytes);
void wvb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
