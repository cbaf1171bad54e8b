// wvb_dsd.cuh -- CUDA kernels and launcher for the DSD decoders of wvb_dsd_core.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "wvb_dsd_core.cuh"

namespace wvb {

// planner class of a DSD block: 0 raw, 3 high, 16+history_bits fast, -1 unknown
inline int dsd_mode_class(const wvb_block_desc &d)
{
    const uint32_t k = d.smem_words;
    switch (k & 15u) {
    case 0: return 0;
    case 3: return 3;
    case 1: return 16 + (int)((k >> 4) & 15u);
    default: return -1;
    }
}

constexpr int DSD_RAW_THREADS = 128;
constexpr int DSD_HIGH_THREADS = 64;
constexpr int DSD_FAST_WARPS = 4;

// The adaptive table of mode 3: one 256-entry table per thread, [256][DSD_HIGH_THREADS] in shared memory (bank == lane).
// It is what limits the kernel to 6 resident warps per SM.  Measured and rejected in round 2: entries packed to 24 bits
// (entry - 0x10000 fits: a 16-bit plane and an 8-bit plane, 768 B per thread, 8-9 warps) with 64 / 96 / 128-thread CTAs:
// 268 / 259 / 267 ms against 254 ms for this layout -- the second load and store per bit and the smaller L1 that the extra
// shared memory leaves cost what the occupancy returns; one-warp CTAs: 363 ms.
constexpr size_t DSD_HIGH_SMEM = (size_t)256 * DSD_HIGH_THREADS * sizeof(int);
struct PtableColumn { // ptable entry i of this thread
    int *base;
    __device__ __forceinline__ explicit PtableColumn(void *smem, int tid) : base((int *)smem + tid) {}
    __device__ __forceinline__ int &operator()(int i) { return base[i * DSD_HIGH_THREADS]; }
};

// mode 0 (DsdUtils.cs:73-82) is a copy plus the crc*3+byte recurrence.  One block per warp: lanes read 4 consecutive
// bytes each (128 B per warp step, coalesced), and the CRC -- an affine recurrence mod 2^32 -- is
// evaluated per 128-byte step as crc = crc*3^128 + sum_l local_l * 3^(124-4l) with a warp reduction.
static __device__ __forceinline__ uint32_t pow3(uint32_t e)
{
    uint32_t r = 1, b = 3;
    while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}

// 16 source bytes starting `mis` bytes into the 32 bytes (a, b): the copy is store-aligned, so the loads are whatever the
// block's position in the file makes them (warp-uniform misalignment: every lane is 16 bytes from its neighbour)
static __device__ __forceinline__ uint4 dsd_extract16(const uint4 a, const uint4 b, uint32_t mis)
{
    uint32_t w0 = a.x, w1 = a.y, w2 = a.z, w3 = a.w, w4 = b.x;
    const uint32_t idx = mis >> 2;
    if (idx == 1) { w0 = a.y; w1 = a.z; w2 = a.w; w3 = b.x; w4 = b.y; }
    else if (idx == 2) { w0 = a.z; w1 = a.w; w2 = b.x; w3 = b.y; w4 = b.z; }
    else if (idx == 3) { w0 = a.w; w1 = b.x; w2 = b.y; w3 = b.z; w4 = b.w; }
    const uint32_t sh = 8u * (mis & 3u);
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// crc*3+byte over the 16 bytes of v, starting from 0: sum b_i * 3^(15-i).  One byte dot product per word (27,9,3,1).
static __device__ __forceinline__ uint32_t dsd_crc16(const uint4 v)
{
    const uint32_t k = 0x0103091bu; // byte 0 (first in memory) x 27, byte 1 x 9, byte 2 x 3, byte 3 x 1
    return ((__dp4a(v.x, k, 0u) * 81u + __dp4a(v.y, k, 0u)) * 81u + __dp4a(v.z, k, 0u)) * 81u + __dp4a(v.w, k, 0u);
}

static __global__ void __launch_bounds__(DSD_RAW_THREADS)
k_dsd_raw(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
          uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results)
{
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * DSD_RAW_THREADS + threadIdx.x) >> 5;
    if (w >= count) return; // whole warp leaves together
    const uint32_t bi = order[w];
    const wvb_block_desc &D = descs[bi];
    if (D.bflags & WVB_BF_MUTE_ALL) { dsd_stale_block(D, out, out_format, &results[bi], lane, 32); return; } // warp-uniform
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD] + 2;
    const uint32_t len = D.sub_len[WVB_SUB_DSD];
    uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
    if (len - 2 < total) total = len - 2; // DsdUtils.cs:77-78
    uint32_t crc = 0xffffffffu;
    // plain byte output of a stereo (or mono) stream into a contiguous frame layout: a copy
    const bool bytes_contig = o.unit == 1 && o.frame_bytes == (uint32_t)o.coded_ch && o.out_ch == o.coded_ch;
    if (bytes_contig) {
        // The tile-movement kernel of the path: 16-byte aligned vector stores, 16-byte aligned vector loads realigned in
        // registers, and the CRC as one byte dot product per word.  A warp step moves 512 bytes; the CRC of a step is
        // crc*3^512 + sum over lanes of crc16(lane) * 3^(16*(31-lane)), one warp reduction.
        uint8_t *dst = o.op;
        uint32_t head = (uint32_t)((16u - ((uintptr_t)dst & 15u)) & 15u);
        if (head > total) head = total;
        for (uint32_t j = 0; j < head; ++j) { // (every lane walks the few head bytes: the CRC stays warp-uniform)
            const uint32_t b = p[j];
            crc = crc * 3u + b;
            if (lane == 0) dst[j] = (uint8_t)(b + (uint32_t)o.add);
        }
        const uint8_t *src = p + head;
        dst += head;
        const uint32_t body = total - head, nchunks = body >> 4;
        const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
        const uint4 *sa = (const uint4 *)(src - mis);
        uint4 *da = (uint4 *)dst;
        const uint32_t flip = o.add ? 0x80808080u : 0u; // +128 per byte
        const uint32_t lane_pow = pow3(16u * (31u - (uint32_t)lane)), step_pow = pow3(512u);
        uint32_t c = 0;
        const uint32_t FULL = 0xffffffffu;
        for (; c + 64 <= nchunks; c += 64) { // two warp steps per trip: both loads are in flight before the first is used
            const uint4 a0 = __ldg(sa + c + lane), a1 = __ldg(sa + c + 32 + lane);
            uint4 e = make_uint4(0, 0, 0, 0);
            if (lane == 31 && mis) e = __ldg(sa + c + 64); // (mis == 0 never looks past its own 16 bytes: no read beyond the payload)
            // the 16 bytes after a lane's own: the next lane's; for lane 31 the first lane's of the next step
            uint4 b0 = make_uint4(__shfl_down_sync(FULL, a0.x, 1), __shfl_down_sync(FULL, a0.y, 1), __shfl_down_sync(FULL, a0.z, 1), __shfl_down_sync(FULL, a0.w, 1));
            uint4 b1 = make_uint4(__shfl_down_sync(FULL, a1.x, 1), __shfl_down_sync(FULL, a1.y, 1), __shfl_down_sync(FULL, a1.z, 1), __shfl_down_sync(FULL, a1.w, 1));
            const uint4 f1 = make_uint4(__shfl_sync(FULL, a1.x, 0), __shfl_sync(FULL, a1.y, 0), __shfl_sync(FULL, a1.z, 0), __shfl_sync(FULL, a1.w, 0));
            if (lane == 31) { b0 = f1; b1 = e; }
            const uint4 v0 = dsd_extract16(a0, b0, mis), v1 = dsd_extract16(a1, b1, mis);
            crc = crc * step_pow + __reduce_add_sync(FULL, dsd_crc16(v0) * lane_pow);
            crc = crc * step_pow + __reduce_add_sync(FULL, dsd_crc16(v1) * lane_pow);
            __stcs(da + c + lane, make_uint4(v0.x ^ flip, v0.y ^ flip, v0.z ^ flip, v0.w ^ flip));
            __stcs(da + c + 32 + lane, make_uint4(v1.x ^ flip, v1.y ^ flip, v1.z ^ flip, v1.w ^ flip));
        }
        for (; c + 32 <= nchunks; c += 32) {
            const uint4 a = __ldg(sa + c + lane);
            uint4 b = make_uint4(__shfl_down_sync(FULL, a.x, 1), __shfl_down_sync(FULL, a.y, 1), __shfl_down_sync(FULL, a.z, 1), __shfl_down_sync(FULL, a.w, 1));
            if (lane == 31 && mis) b = __ldg(sa + c + 32);
            const uint4 v = dsd_extract16(a, b, mis);
            crc = crc * step_pow + __reduce_add_sync(FULL, dsd_crc16(v) * lane_pow);
            __stcs(da + c + lane, make_uint4(v.x ^ flip, v.y ^ flip, v.z ^ flip, v.w ^ flip));
        }
        const uint32_t rest = nchunks - c; // a last, partial warp step
        if (rest) {
            const bool mine = (uint32_t)lane < rest;
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (mine) {
                a = __ldg(sa + c + lane);
                if (mis) b = __ldg(sa + c + lane + 1);
            }
            const uint4 v = dsd_extract16(a, b, mis);
            const uint32_t term = mine ? dsd_crc16(v) * pow3(16u * (rest - 1u - (uint32_t)lane)) : 0u;
            crc = crc * pow3(16u * rest) + __reduce_add_sync(0xffffffffu, term);
            if (mine) __stcs(da + c + lane, make_uint4(v.x ^ flip, v.y ^ flip, v.z ^ flip, v.w ^ flip));
        }
        for (uint32_t j = head + (nchunks << 4); j < total; ++j) { // fewer than 16 bytes
            const uint32_t b = p[j];
            crc = crc * 3u + b;
            if (lane == 0) o.op[j] = (uint8_t)(b + (uint32_t)o.add);
        }
        if (lane == 0) dsd_finish(D, &results[bi], (int)crc, false, 0, 0);
        return;
    }
    // any other layout (int32 output, FALSE_STEREO duplication, channel slots inside wider frames): 4 values per lane and step
    const uint32_t lane_pow = pow3(124u - 4u * (uint32_t)lane), step_pow = pow3(128u);
    const uint32_t full = total & ~127u;
    for (uint32_t base = 0; base < full; base += 128) {
        const uint32_t j = base + 4u * (uint32_t)lane;
        const uint32_t b0 = p[j], b1 = p[j + 1], b2 = p[j + 2], b3 = p[j + 3];
        const uint32_t local = ((b0 * 3u + b1) * 3u + b2) * 3u + b3;
        crc = crc * step_pow + __reduce_add_sync(0xffffffffu, local * lane_pow);
        o.put(j, (int)b0); o.put(j + 1, (int)b1); o.put(j + 2, (int)b2); o.put(j + 3, (int)b3);
    }
    if (lane == 0) { // tail (< 128 values) and the verdict
        for (uint32_t j = full; j < total; ++j) {
            const uint32_t b = p[j];
            crc = crc * 3u + b;
            o.put(j, (int)b);
        }
        dsd_finish(D, &results[bi], (int)crc, false, 0, 0);
    }
}

static __global__ void __launch_bounds__(DSD_HIGH_THREADS)
k_dsd_high(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
           uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results, const int *__restrict__ ptables)
{
    extern __shared__ int dsd_smem[];
    const uint32_t i = blockIdx.x * DSD_HIGH_THREADS + threadIdx.x;
    const bool valid = i < count;
    const uint32_t bi = order[valid ? i : count - 1];
    const wvb_block_desc &D = descs[bi];
    PtableColumn PT(dsd_smem, (int)threadIdx.x);
    dsd_decode_high(PT, ptables + 256 * dsd_key_rate(D.smem_words), in, D, out, out_format, &results[bi], valid);
}

// ---- mode 1 ("fast"), two kernels -------------------------------------------------------------
// (1) k_dsd_fast_build, one block per warp: RLE-decode the probability tables and prefix-sum them into a global scratch
//     table (bins x 256 u16 cumulative sums, 512 B per history bin).
// (2) k_dsd_fast_dec, one block per THREAD: the range decoder.  A warp-per-block decoder spends ~130 warp instructions per
//     symbol on uniform work (ncu: issue-bound at 73%, profiles/r01_ncu_dsd_fast_warp_per_block.txt); per thread the same
//     work advances 32 blocks.  Each thread keeps an 8-entry coarse index per bin (every 32nd cumulative sum) in shared
//     memory [word][thread]; the symbol lookup is a coarse count there plus a fine count over one 64-byte cell of its
//     global table (one DRAM access).  The kernel is memory-latency bound, so the small index (16 B per bin) buys occupancy.
constexpr int DSD_FAST_DEC_THREADS = 128;
constexpr size_t DSD_FAST_TABLE_STRIDE = 32 * 512; // scratch bytes per block (room for 32 bins)

struct DsdFastMeta { uint32_t at; uint32_t bins; }; // payload offset after the tables (0 = rejected), history bins

static __global__ void __launch_bounds__(DSD_FAST_WARPS * 32)
k_dsd_fast_build(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
                 uint8_t *__restrict__ scratch, DsdFastMeta *__restrict__ meta)
{
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * DSD_FAST_WARPS + wic;
    if (w >= count) return; // whole warp leaves together
    const wvb_block_desc &D = descs[order[w]];
    DsdFastTables T;
    T.summed = (uint16_t *)(scratch + (size_t)w * DSD_FAST_TABLE_STRIDE);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
    int bins = 1;
    const uint32_t at = dsd_fast_build(T, p, D.sub_len[WVB_SUB_DSD], lane, 32, bins);
    __syncwarp();
    if (at != 0) {
        dsd_fast_sums(T, bins, lane, 32, [&](uint32_t local) {
            uint32_t incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            return incl - local;
        });
    }
    if (lane == 0) { meta[w].at = at; meta[w].bins = (uint32_t)bins; }
}

struct CoarseColumn { // word i of this thread's coarse index: [word][DSD_FAST_DEC_THREADS]
    uint32_t *base;
    __device__ __forceinline__ uint32_t &operator()(int i) { return base[i * DSD_FAST_DEC_THREADS]; }
};

// CW = coarse-index words per history bin: 8 (every 16th entry, 32-byte cells) or 4 (every 32nd entry, 64-byte cells).
// Measured on B200, 160k blocks: 16 bins  CW=8 224 ms / CW=4 362 ms (more resident blocks thrash L2);  32 bins  CW=8 353 ms / CW=4 138 ms.
template <int CW>
static __global__ void __launch_bounds__(DSD_FAST_DEC_THREADS)
k_dsd_fast_dec(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
               uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results, const uint8_t *__restrict__ scratch,
               const DsdFastMeta *__restrict__ meta)
{
    extern __shared__ int dsd_smem[];
    const uint32_t i = blockIdx.x * DSD_FAST_DEC_THREADS + threadIdx.x;
    const bool valid = i < count; // lanes past the end stay in the warp with no work (warp-synchronous loop)
    const uint32_t slot = valid ? i : count - 1;
    const uint32_t bi = order[slot];
    const wvb_block_desc &D = descs[bi];
    const DsdFastMeta M = meta[slot];
    const bool good = valid && M.at != 0;
    const int bins = good ? (int)M.bins : 1;
    DsdFastTables T;
    T.summed = (uint16_t *)(scratch + (size_t)slot * DSD_FAST_TABLE_STRIDE);
    CoarseColumn CO{(uint32_t *)dsd_smem + threadIdx.x};
    if (good)
        for (int b = 0; b < bins; ++b)
            for (int k = 0; k < CW; ++k) { // two coarse entries per word: the last entries of cells 2k and 2k+1
                constexpr int CELL = 128 / CW; // 16 or 32 table entries per cell
                CO(b * CW + k) = (uint32_t)T.summed[b * 256 + 2 * CELL * k + CELL - 1] | ((uint32_t)T.summed[b * 256 + 2 * CELL * k + 2 * CELL - 1] << 16);
            }
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const bool mono = o.coded_ch == 1;
    const uint32_t total = good ? D.block_samples * (uint32_t)o.coded_ch : 0;
    const bool packed = o.unit == 1 && o.frame_bytes == (uint32_t)o.coded_ch && o.out_ch == o.coded_ch && (((uintptr_t)o.op) & 3) == 0;
    const uint32_t addv = (uint32_t)o.add;
    uint32_t acc = 0;
    int crc = -1;
    bool failed = false;
    uint32_t fail_at = total;
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
    dsd_fast_decode(T, bins, p, D.sub_len[WVB_SUB_DSD], good ? M.at : 6u, mono, total,
        [&](const uint16_t *, int p0) { return CO(p0 * CW + CW - 1) >> 16; }, // row[255] is the last coarse entry
        [&](const uint16_t *row, int p0, uint32_t index, uint32_t &below, uint32_t &cur) {
            const uint32_t idx2 = index | (index << 16);
            int c = 0;
#pragma unroll
            for (int k = 0; k < CW; ++k) c += __popc(__vcmpleu2(CO(p0 * CW + k), idx2));
            c >>= 4; // coarse cell (index < row[255] keeps it inside the row)
            constexpr int CELL = 128 / CW;
            const uint4 *q = (const uint4 *)(row + CELL * c); // the cell: 16 or 32 cumulative sums, contiguous in this block's table
            int f = 0;
#pragma unroll
            for (int k = 0; k < CELL / 8; ++k) {
                const uint4 v = q[k];
                f += __popc(__vcmpleu2(v.x, idx2)) + __popc(__vcmpleu2(v.y, idx2)) + __popc(__vcmpleu2(v.z, idx2)) + __popc(__vcmpleu2(v.w, idx2));
            }
            const int code = CELL * c + (f >> 4);
            cur = row[code]; // same lines as the cell just read
            below = code > 0 ? row[code - 1] : 0u;
            return code;
        },
        [&](uint32_t j, int code) {
            if (packed) { // four byte values per 32-bit store
                acc |= (((uint32_t)code + addv) & 0xffu) << (8 * (j & 3u));
                if ((j & 3u) == 3u) { *(uint32_t *)(o.op + (j & ~3u)) = acc; acc = 0; }
            } else
                o.put(j, code);
        },
        [&](uint32_t delivered) {
            if (packed)
                for (uint32_t k = delivered & ~3u; k < delivered; ++k) o.op[k] = (uint8_t)(acc >> (8 * (k & 3u)));
        },
        crc, failed, fail_at);
    if (valid) {
        if (!good) { results[bi].crc = -1; results[bi].crc_x = -1; results[bi].mute_from = 0; results[bi].rflags = WVB_RF_BAD_BLOCK | WVB_RF_MUTED | WVB_RF_CRC_ERROR; }
        else dsd_finish(D, &results[bi], crc, failed, mono ? fail_at : fail_at >> 1, 0);
    }
}

// second pass over the DSD blocks of a launch: 0x55 fill for muted pieces (DsdUtils.cs:104-117), after every decode
// thread of the batch is done because the fill can land in the previous block's output (it starts at call-buffer index 0)
static __global__ void k_dsd_mute_fix(const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
                                      uint8_t *__restrict__ out, int out_format, const wvb_block_result *__restrict__ results)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t bi = order[i];
    const wvb_block_result &R = results[bi];
    if (!(R.rflags & WVB_RF_MUTED)) return;
    const wvb_block_desc &D = descs[bi];
    const int unit = out_format == WVB_OUT_INT32 ? 4 : 1, add = out_format == WVB_OUT_PCM ? 128 : 0;
    const uint32_t frame_bytes = (uint32_t)unit * D.out_stride;
    const uint32_t n = D.block_samples;
    uint32_t ps = R.mute_from;
    while (ps < n) {
        uint32_t p0, pe;
        piece_of(D, n, ps, p0, pe);
        // the fill starts at index 0 of the caller's buffer for that call, not at the piece (quirk C-11)
        const int64_t start = (int64_t)ps - (int64_t)call_lookback(D, ps);
        uint8_t *q = out + D.out_offset + start * (int64_t)frame_bytes;
        for (uint32_t k = 0; k < pe - ps; ++k, q += frame_bytes)
            for (int c = 0; c < D.out_stride; ++c) store_unit(q + c * unit, 0x55, unit, add);
        ps = pe;
    }
}

struct DsdDeviceTables {
    std::mutex mu;
    int *d_ptables[64] = {nullptr};
};
inline DsdDeviceTables &dsd_tables() { static DsdDeviceTables t; return t; }

inline const int *dsd_device_ptables(int device)
{
    DsdDeviceTables &T = dsd_tables();
    std::lock_guard<std::mutex> g(T.mu);
    if (device < 0 || device >= 64) return nullptr;
    if (!T.d_ptables[device]) {
        std::vector<int> host(256 * 256);
        for (int r = 0; r < 256; r++) dsd_init_ptable_host(host.data() + 256 * r, r, 20);
        int *d = nullptr;
        if (cudaMalloc((void **)&d, host.size() * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemcpy(d, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        T.d_ptables[device] = d;
    }
    return T.d_ptables[device];
}

// returns WVB_OK or an error code; the caller reports cudaGetLastError text
inline int launch_dsd(int cls, const uint8_t *din, const wvb_block_desc *d_descs, const uint32_t *d_order, uint32_t count, uint8_t *dout,
                      int out_format, wvb_block_result *dres, cudaStream_t s, size_t smem_optin, int device, int *launches,
                      uint8_t *scratch, uint8_t *scratch_meta, uint32_t scratch_slot0)
{
    if (count == 0) return WVB_OK;
    if (cls == 0) {
        k_dsd_raw<<<(count + DSD_RAW_THREADS / 32 - 1) / (DSD_RAW_THREADS / 32), DSD_RAW_THREADS, 0, s>>>(din, d_descs, d_order, count, dout, out_format, dres);
    } else if (cls == 3) {
        const int *pt = dsd_device_ptables(device);
        if (!pt) return WVB_E_CUDA;
        const size_t smem = DSD_HIGH_SMEM;
        if (smem > smem_optin) return WVB_E_ARG;
        if (cudaFuncSetAttribute((const void *)k_dsd_high, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WVB_E_CUDA;
        k_dsd_high<<<(count + DSD_HIGH_THREADS - 1) / DSD_HIGH_THREADS, DSD_HIGH_THREADS, smem, s>>>(din, d_descs, d_order, count, dout, out_format, dres, pt);
    } else if (cls >= 16 && cls <= 16 + 5) {
        if (!scratch) return WVB_E_ARG;
        const int bins = 1 << (cls - 16);
        uint8_t *tab = scratch + (size_t)scratch_slot0 * DSD_FAST_TABLE_STRIDE;
        DsdFastMeta *meta = (DsdFastMeta *)(scratch_meta) + scratch_slot0;
        k_dsd_fast_build<<<(count + DSD_FAST_WARPS - 1) / DSD_FAST_WARPS, DSD_FAST_WARPS * 32, 0, s>>>(din, d_descs, d_order, count, tab, meta);
        if (cudaGetLastError() != cudaSuccess) return WVB_E_CUDA;
        if (launches) (*launches)++;
        const int cw = bins >= 32 ? 4 : 8;
        const size_t smem = (size_t)bins * cw * sizeof(uint32_t) * DSD_FAST_DEC_THREADS;
        if (smem > smem_optin) return WVB_E_ARG;
        auto kern = cw == 4 ? k_dsd_fast_dec<4> : k_dsd_fast_dec<8>;
        if (cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WVB_E_CUDA;
        kern<<<(count + DSD_FAST_DEC_THREADS - 1) / DSD_FAST_DEC_THREADS, DSD_FAST_DEC_THREADS, smem, s>>>(din, d_descs, d_order, count, dout, out_format, dres,
                                                                                                    tab, meta);
    } else
        return WVB_E_ARG;
    if (cudaGetLastError() != cudaSuccess) return WVB_E_CUDA;
    if (launches) (*launches)++;
    return WVB_OK;
}

inline int launch_dsd_mute_fix(const wvb_block_desc *d_descs, const uint32_t *d_order, uint32_t count, uint8_t *dout, int out_format,
                               const wvb_block_result *dres, cudaStream_t s, int *launches)
{
    if (count == 0) return WVB_OK;
    k_dsd_mute_fix<<<(count + 127) / 128, 128, 0, s>>>(d_descs, d_order, count, dout, out_format, dres);
    if (cudaGetLastError() != cudaSuccess) return WVB_E_CUDA;
    if (launches) (*launches)++;
    return WVB_OK;
}

} // namespace wvb
