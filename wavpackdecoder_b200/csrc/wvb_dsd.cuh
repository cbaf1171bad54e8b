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

struct PtableColumn { // ptable entry i of this thread: [256][DSD_HIGH_THREADS]
    int *base;
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

static __global__ void __launch_bounds__(DSD_RAW_THREADS)
k_dsd_raw(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
          uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results)
{
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * DSD_RAW_THREADS + threadIdx.x) >> 5;
    if (w >= count) return; // whole warp leaves together
    const uint32_t bi = order[w];
    const wvb_block_desc &D = descs[bi];
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD] + 2;
    const uint32_t len = D.sub_len[WVB_SUB_DSD];
    uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
    if (len - 2 < total) total = len - 2; // DsdUtils.cs:77-78
    const uint32_t lane_pow = pow3(124u - 4u * (uint32_t)lane), step_pow = pow3(128u);
    uint32_t crc = 0xffffffffu;
    const uint32_t full = total & ~127u;
    // fast path: plain byte output of a stereo (or mono) stream into a contiguous frame layout
    const bool bytes_contig = o.unit == 1 && o.frame_bytes == (uint32_t)o.coded_ch && o.out_ch == o.coded_ch;
    for (uint32_t base = 0; base < full; base += 128) {
        const uint32_t j = base + 4u * (uint32_t)lane;
        const uint32_t b0 = p[j], b1 = p[j + 1], b2 = p[j + 2], b3 = p[j + 3];
        const uint32_t local = ((b0 * 3u + b1) * 3u + b2) * 3u + b3;
        crc = crc * step_pow + __reduce_add_sync(0xffffffffu, local * lane_pow);
        if (bytes_contig) {
            uint8_t *q = o.op + j;
            const uint32_t a = (uint32_t)o.add;
            q[0] = (uint8_t)(b0 + a); q[1] = (uint8_t)(b1 + a); q[2] = (uint8_t)(b2 + a); q[3] = (uint8_t)(b3 + a);
        } else {
            o.put(j, (int)b0); o.put(j + 1, (int)b1); o.put(j + 2, (int)b2); o.put(j + 3, (int)b3);
        }
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
    PtableColumn PT{dsd_smem + threadIdx.x};
    dsd_decode_high(PT, ptables + 256 * dsd_key_rate(D.smem_words), in, D, out, out_format, &results[bi], valid);
}

static __global__ void __launch_bounds__(DSD_FAST_WARPS * 32)
k_dsd_fast(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
           uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results, int bins_max)
{
    extern __shared__ int dsd_smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * DSD_FAST_WARPS + wic;
    if (w >= count) return; // whole warp leaves together
    const uint32_t bi = order[w];
    const wvb_block_desc &D = descs[bi];
    DsdFastTables T;
    T.summed = (uint16_t *)((uint8_t *)dsd_smem + (size_t)wic * bins_max * 512);
    const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
    const uint32_t len = D.sub_len[WVB_SUB_DSD];
    int bins = 1;
    const uint32_t at = dsd_fast_build(T, p, len, lane, 32, bins);
    __syncwarp();
    if (at == 0 || bins > bins_max) {
        if (lane == 0) { results[bi].crc = -1; results[bi].crc_x = -1; results[bi].mute_from = 0; results[bi].rflags = WVB_RF_BAD_BLOCK | WVB_RF_MUTED | WVB_RF_CRC_ERROR; }
        return;
    }
    dsd_fast_sums(T, bins, lane, 32, [&](uint32_t local) {
        uint32_t incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        return incl - local;
    });
    __syncwarp();
    DsdOut o;
    dsd_out_init(o, D, out, out_format);
    const bool mono = o.coded_ch == 1;
    const uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
    int crc = -1;
    bool failed = false;
    uint32_t fail_at = total;
    // per-bin reciprocal of summed[bin][255] (16-bit divisor d, 32-bit numerator n): with m = ceil(2^48 / d),
    // n / d == (n * m) >> 48 exactly (m*d - 2^48 < d <= 2^16, Granlund-Montgomery).  Lane b owns bin b.
    uint64_t my_recip = 0;
    if (lane < bins) {
        const uint32_t d = T.summed[lane * 256 + 255];
        my_recip = d ? (((uint64_t)1 << 48) + d - 1) / d : 0;
    }
    int mine = 0; // the value this lane will store at the next flush (lane == j & 31)
    dsd_fast_decode(T, bins, p, len, at, mono, total,
        [&](const uint16_t *row, uint32_t index) {
            const uint4 v = *(const uint4 *)(row + lane * 8);
            const uint32_t idx2 = index | (index << 16);
            const int c = (__popc(__vcmpleu2(v.x, idx2)) + __popc(__vcmpleu2(v.y, idx2)) + __popc(__vcmpleu2(v.z, idx2)) + __popc(__vcmpleu2(v.w, idx2))) >> 4;
            return (int)__reduce_add_sync(0xffffffffu, (unsigned)c);
        },
        [&](int p0, uint32_t n) {
            const uint64_t m = __shfl_sync(0xffffffffu, my_recip, p0);
            return (uint32_t)__umul64hi((uint64_t)n << 16, m);
        },
        [&](uint32_t j, int code) {
            if ((int)(j & 31u) == lane) mine = code;
            if ((j & 31u) == 31u) o.put((j & ~31u) + (uint32_t)lane, mine); // one coalesced store per 32 values
        },
        [&](uint32_t count) { // values of the last, partial group
            if ((count & 31u) != 0 && (uint32_t)lane < (count & 31u)) o.put((count & ~31u) + (uint32_t)lane, mine);
        },
        crc, failed, fail_at);
    if (lane == 0) dsd_finish(D, &results[bi], crc, failed, mono ? fail_at : fail_at >> 1, 0);
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
    const uint32_t n = D.block_samples, chunk = D.chunk_samples ? D.chunk_samples : 0xffffffffu;
    uint32_t first_len = D.chunk_first < n ? D.chunk_first : n;
    if (first_len == 0) first_len = chunk < n ? chunk : n;
    uint32_t ps = R.mute_from;
    while (ps < n) {
        const uint32_t pe = ps < first_len ? first_len : (ps + chunk < n ? ps + chunk : n);
        // the fill starts at index 0 of the caller's buffer for that call, not at the piece (quirk C-11)
        int64_t start = ps;
        if (ps == 0 && D.chunk_first != 0 && D.chunk_first < chunk) start = -(int64_t)(chunk - D.chunk_first);
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
                      int out_format, wvb_block_result *dres, cudaStream_t s, size_t smem_optin, int device, int *launches)
{
    if (count == 0) return WVB_OK;
    if (cls == 0) {
        k_dsd_raw<<<(count + DSD_RAW_THREADS / 32 - 1) / (DSD_RAW_THREADS / 32), DSD_RAW_THREADS, 0, s>>>(din, d_descs, d_order, count, dout, out_format, dres);
    } else if (cls == 3) {
        const int *pt = dsd_device_ptables(device);
        if (!pt) return WVB_E_CUDA;
        const size_t smem = (size_t)256 * DSD_HIGH_THREADS * sizeof(int);
        if (smem > smem_optin) return WVB_E_ARG;
        if (cudaFuncSetAttribute((const void *)k_dsd_high, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WVB_E_CUDA;
        k_dsd_high<<<(count + DSD_HIGH_THREADS - 1) / DSD_HIGH_THREADS, DSD_HIGH_THREADS, smem, s>>>(din, d_descs, d_order, count, dout, out_format, dres, pt);
    } else if (cls >= 16 && cls <= 16 + 5) {
        const int bins = 1 << (cls - 16);
        const size_t smem = (size_t)DSD_FAST_WARPS * bins * 512;
        if (smem > smem_optin) return WVB_E_ARG;
        if (cudaFuncSetAttribute((const void *)k_dsd_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return WVB_E_CUDA;
        k_dsd_fast<<<(count + DSD_FAST_WARPS - 1) / DSD_FAST_WARPS, DSD_FAST_WARPS * 32, smem, s>>>(din, d_descs, d_order, count, dout, out_format, dres, bins);
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
