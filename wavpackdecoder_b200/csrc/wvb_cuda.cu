// wvb_cuda.cu -- CUDA side of libwvb: kernels, launch planning, the wvb_batch_* C ABI.
//
// Execution model: WavPack blocks are independent, each one a strictly serial
// chain (adaptive Golomb words -> adaptive decorrelation).  One block is decoded by one
// THREAD; a warp decodes 32 blocks in lock step.  The planner sorts blocks so that a warp
// holds blocks with the same kernel variant, the same decorrelation term list and the same
// length, which keeps the warp convergent.  Per-thread decorrelation state lives in shared
// memory laid out [slot][thread] (bank == lane for any per-lane slot index).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "wvb_checksum.cuh"
#include "wvb_dsd.cuh"
#include "wvb_dsf.cuh"
#include "wvb_md5.cuh"
#include "wvb_pcm.cuh"
#include "wvb_plan.h"

namespace {

#ifndef WVB_CTA
#define WVB_CTA 32 // threads per decode CTA: one warp (measured against 64 / 128: 76.9 / 77.1 / 77.8 ms on the bench launch, see k_decode_pcm)
#endif
#ifndef WVB_STAGE_OUTPUT
#define WVB_STAGE_OUTPUT 1 // 0: experiment builds without the shared-memory output staging of the 16-bit stereo kernels
#endif
constexpr int CTA_THREADS = WVB_CTA;

thread_local std::string g_last_error;
int set_error(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) return set_error(WVB_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

struct Launch { int variant; int cls; uint32_t first, count; };

// Dynamic shared memory of a decode CTA: cls words per thread laid out [slot][thread] (a thread's "column": bank == lane
// for any per-lane slot index), then, for the kernels that stage their output (16-bit stereo PCM, wvb_pcm.cuh Stage16),
// 8 bytes of staging bookkeeping per thread.  RESERVE: column slots at the top kept for the staging ring (the generic
// decorrelator's state lives in the column during the sample loop; the in-register kernels' state is dead by then and the
// ring reuses slots 0..15).
template <int CTA, bool STAGED = false, int RESERVE = 0> struct SharedColumn {
    static constexpr int kThreads = CTA;
    static constexpr bool kStaged = STAGED;
    int *base;   // this thread's column
    int *origin; // word 0 of the CTA's dynamic shared memory
    __device__ __forceinline__ int &operator()(int i) { return base[i * CTA]; }
    __device__ __forceinline__ int slots() const // column slots this launch's dynamic shared memory provides
    {
        uint32_t bytes;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(bytes));
        return (int)((bytes - (STAGED ? CTA * 8u : 0u)) / (CTA * sizeof(int)));
    }
    __device__ __forceinline__ int cap() const { return slots() - RESERVE; } // slots for decorrelation state
    __device__ __forceinline__ int ring_slot0() const { return RESERVE ? cap() : 0; }
    __device__ __forceinline__ uint2 *stage_meta() const { return (uint2 *)(origin + slots() * CTA); } // [thread]
};

// CTA: threads per CTA.  Nothing in the decoder is CTA-wide (no barrier, no shared data between threads), so the CTA size
// only sets the granularity at which shared memory and registers are handed out and at which a launch's tail drains.
// One-warp CTAs are the default: blocks with long term lists (150-250 words of decorrelation state per thread) fit 10 warps
// per SM that way where 128-thread CTAs fit 8; for the 88-register kernels (20 warps either way) the gain is the finer launch tail.
constexpr int CTA_SMALL = 32, CTA_FIXED_D = 64;
// MINB: register cap, as the number of 128-thread CTAs' worth of threads that must fit an SM (0: no cap)
template <bool STEREO, bool HYB, bool GENFIX, class DEC, int MINB = 0, bool F16 = false, int CTA = CTA_THREADS>
__global__ void __launch_bounds__(CTA, MINB * 128 / CTA)
k_decode_pcm(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order,
             uint32_t count, uint8_t *__restrict__ out, int out_format, wvb_block_result *__restrict__ results, uint32_t spread)
{
    extern __shared__ int smem[];
    // spread (a power-of-two exponent, 0..5): a launch too small to fill the GPU gives every block a group of 2^spread lanes
    // and lets only the group's last lane work -- see pick_spread()
    const uint32_t i = blockIdx.x * CTA + threadIdx.x, slot = i >> spread, gmask = (1u << spread) - 1u;
    const bool valid = slot < count && (i & gmask) == gmask; // lanes without work keep running: the decode loop is warp-synchronous
    const uint32_t bi = order[slot < count ? slot : count - 1];
    constexpr bool staged = F16 && STEREO && !GENFIX && WVB_STAGE_OUTPUT;
    using Column = SharedColumn<CTA, staged, (staged && !DEC::kFixed) ? wvb::STAGE_RING_SLOTS : 0>;
    Column SM{smem + threadIdx.x, smem};
    wvb::decode_block_pcm<STEREO, HYB, GENFIX, Column, DEC, F16>(SM, in, descs[bi], out, out_format, &results[bi], valid);
}

using GenS = wvb::GenericDecorr<true>;
using GenM = wvb::GenericDecorr<false>;
using FixS = wvb::FixedDecorr<true, WVB_FIXED_STEREO_TERMS>;
using FixM = wvb::FixedDecorr<false, WVB_FIXED_MONO_TERMS>;
using FixSB = wvb::FixedDecorr<true, WVB_FIXED_STEREO_B_TERMS>;
using FixSC = wvb::FixedDecorr<true, WVB_FIXED_STEREO_C_TERMS>;
using FixSD = wvb::FixedDecorr<true, WVB_FIXED_STEREO_D_TERMS>;

} // namespace

struct wvb_batch {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; // h2d0,h2d1(k0),k1,d2h1
    uint8_t *d_in = nullptr; size_t d_in_cap = 0;
    uint8_t *d_out = nullptr; size_t d_out_cap = 0;
    uint64_t *d_md5_ranges = nullptr; size_t d_md5_ranges_cap = 0;
    uint8_t *d_md5_out = nullptr; size_t d_md5_out_cap = 0;
    wvb_block_desc *d_descs = nullptr; size_t d_descs_cap = 0;
    uint32_t *d_order = nullptr; size_t d_order_cap = 0;
    wvb_block_result *d_results = nullptr; size_t d_results_cap = 0;
    uint8_t *d_scratch = nullptr; size_t d_scratch_cap = 0;       // DSD fast-mode tables, one 16 KB slot per table position
    uint8_t *d_scratch_meta = nullptr; size_t d_scratch_meta_cap = 0;
    std::vector<uint32_t> order;
    std::vector<Launch> plan;
    size_t prepared_n = 0; bool prepared = false; int prepared_fmt = -1;
    uint64_t prepared_in_extent = 0, prepared_out_extent = 0; // slab sizes the prepared table needs
    // pinned staging of the small per-call arrays: copies from / into pageable memory are staged by the driver and block the
    // host thread on the stream they are queued on
    uint32_t *h_order = nullptr; size_t h_order_cap = 0;
    wvb_block_desc *h_descs = nullptr; size_t h_descs_cap = 0; // table slices of wvb_batch_decode_files
    wvb_block_result *h_results = nullptr; size_t h_results_cap = 0;
    wvb_block_result *pending_results = nullptr; size_t pending_n = 0; bool pending_copy = false;
    // WVB_TRACE=1: per-segment timeline of the pipelined host-buffer decode, printed by wvb_batch_wait
    struct TraceSeg { cudaEvent_t up, dec, down; size_t blocks; uint64_t in_bytes, out_bytes; double host_ms; };
    std::vector<TraceSeg> trace;
    cudaEvent_t trace_t0 = nullptr;
    double trace_host_total_ms = 0;
    int launches = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr; // copy streams of the pipelined host-buffer path
    std::vector<cudaEvent_t> seg_ev;
    std::vector<cudaStream_t> seg_streams; // kernels of different segments run concurrently (a block's decode time is a
                                           // serial-chain latency, so small launches do not finish sooner than big ones)
    int sm_count = 148;
    size_t smem_optin = 0;
    bool timed = false;
};

namespace {

template <class T> int ensure_pinned(T *&p, size_t &cap, size_t need)
{
    if (need <= cap) return WVB_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    const size_t want = need + need / 8 + 256;
    CUDA_TRY(cudaHostAlloc((void **)&p, want * sizeof(T), cudaHostAllocDefault));
    cap = want;
    return WVB_OK;
}

template <class T> int ensure(T *&p, size_t &cap, size_t need)
{
    if (need <= cap) return WVB_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = need + need / 8 + 256;
    CUDA_TRY(cudaMalloc((void **)&p, want * sizeof(T)));
    cap = want;
    return WVB_OK;
}

typedef void (*pcm_kernel_t)(const uint8_t *, const wvb_block_desc *, const uint32_t *, uint32_t, uint8_t *, int, wvb_block_result *, uint32_t);

// the one-warp-CTA builds of the generic kernels (large shared-memory classes)
pcm_kernel_t pcm_kernel_small(int variant)
{
    switch (variant) {
    case wvb::V_MONO: return k_decode_pcm<false, false, false, GenM, 0, false, CTA_SMALL>;
    case wvb::V_STEREO: return k_decode_pcm<true, false, false, GenS, 0, false, CTA_SMALL>;
    case wvb::V_STEREO | wvb::V_F16: return k_decode_pcm<true, false, false, GenS, 0, true, CTA_SMALL>;
    case wvb::V_MONO | wvb::V_GENFIX: return k_decode_pcm<false, false, true, GenM, 0, false, CTA_SMALL>;
    case wvb::V_STEREO | wvb::V_GENFIX: return k_decode_pcm<true, false, true, GenS, 0, false, CTA_SMALL>;
    case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID: return k_decode_pcm<false, true, true, GenM, 0, false, CTA_SMALL>;
    case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID: return k_decode_pcm<true, true, true, GenS, 0, false, CTA_SMALL>;
    default: return nullptr;
    }
}

pcm_kernel_t pcm_kernel(int variant)
{
    switch (variant) {
    case wvb::V_MONO: return k_decode_pcm<false, false, false, GenM>;
    case wvb::V_STEREO: return k_decode_pcm<true, false, false, GenS>;
    case wvb::V_STEREO | wvb::V_F16: return k_decode_pcm<true, false, false, GenS, 0, true>;
    case wvb::V_MONO | wvb::V_FIXED: return k_decode_pcm<false, false, false, FixM>;
    case wvb::V_STEREO | wvb::V_FIXED: return k_decode_pcm<true, false, false, FixS>;
    // (Round 1 also shipped builds of the two five-term kernels capped at 80 registers, 6 CTAs per SM with ~40 B of spills, and
    // used them for launches of more than one wave.  With the round-2 decoder the uncapped build -- 88 registers, 5 CTAs, no
    // spills -- wins at every size: 75.7 vs 76.6 ms at 200k blocks, 50.5 vs 54.4 ms at 120k; with the staged output the
    // capped build spills the flush: 77.0 vs 78.8 ms.  The capped builds are gone.)
    case wvb::V_STEREO | wvb::V_FIXED | wvb::V_F16: return k_decode_pcm<true, false, false, FixS, 0, true>;
    case wvb::V_STEREO | wvb::V_FIXED_B: return k_decode_pcm<true, false, false, FixSB>;
    case wvb::V_STEREO | wvb::V_FIXED_B | wvb::V_F16: return k_decode_pcm<true, false, false, FixSB, 0, true>;
    case wvb::V_STEREO | wvb::V_FIXED_C: return k_decode_pcm<true, false, false, FixSC>;
    case wvb::V_STEREO | wvb::V_FIXED_C | wvb::V_F16: return k_decode_pcm<true, false, false, FixSC, 0, true>;
    // the 16-term list: 228-255 registers, i.e. 256 resident threads per SM whatever the CTA size; in 64-thread CTAs the
    // launch tail is finer (76.5 vs 79.0 ms per 60 000 blocks).  Capping the registers for a fifth CTA spills the history: 244 ms.
    case wvb::V_STEREO | wvb::V_FIXED_D: return k_decode_pcm<true, false, false, FixSD, 2, false, CTA_FIXED_D>;
    case wvb::V_STEREO | wvb::V_FIXED_D | wvb::V_F16: return k_decode_pcm<true, false, false, FixSD, 2, true, CTA_FIXED_D>;
    case wvb::V_MONO | wvb::V_GENFIX: return k_decode_pcm<false, false, true, GenM>;
    case wvb::V_STEREO | wvb::V_GENFIX: return k_decode_pcm<true, false, true, GenS>;
    case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID: return k_decode_pcm<false, true, true, GenM>;
    case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID: return k_decode_pcm<true, true, true, GenS>;
    // float / int32 / hybrid blocks with the stock term lists: the same fixup and entropy code over in-register decorrelation
    // (float 89.6 -> 64.6 ms, int32 + WVX 107.5 -> 80.5 ms per 120 000 blocks).  The hybrid kernels carry more state and pay
    // for the registers in resident CTAs; measured per register cap (generic / uncapped / 4 CTAs / 5 CTAs): stereo, 160 000
    // blocks: 159.9 / 154 (138 registers) / 155.1 (128) / 214.6 ms (96, 292 B of spills); mono, 240 000 blocks:
    // 114.5 / 104.6 (114) / 103.2 (118) / 94.6 ms (96, 44 B of spills).
    case wvb::V_MONO | wvb::V_GENFIX | wvb::V_FIXED: return k_decode_pcm<false, false, true, FixM>;
    case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_FIXED: return k_decode_pcm<true, false, true, FixS>;
    case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID | wvb::V_FIXED: return k_decode_pcm<false, true, true, FixM, 5>;
    case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID | wvb::V_FIXED: return k_decode_pcm<true, true, true, FixS, 4>;
    default: return nullptr;
    }
}

// Small launches.  A block is a serial chain, so a launch of a few hundred blocks (BASELINE configs[0]: one 60 s file, 120
// blocks) cannot use the machine whatever the mapping, and its duration is the latency of one chain: 17.6 ms for a 22 050-
// sample stereo block, the same from 320 to 18 000 blocks per launch (measured, round 2).  Packed 32 to a warp every lane
// pays for every branch direction any of the 32 takes; alone in its warp a lane executes its own path only.  That is worth
// 10 % (15.96 ms at 320 blocks) -- the chain is bound by the dependent-issue latency of its ~300 instructions per frame, not
// by divergence -- and only while every block can have a scheduler to itself: with 2 to 16 blocks per warp over more warps
// the same launches got 3-9 % slower (2000 / 10 000 / 18 000 blocks: 19.5 / 20.8 / 21.1 vs 18.8 / 19.4 / 19.4 ms).
// So: one lane per warp (spread = 5) up to one block per scheduler, packed warps otherwise.
uint32_t pick_spread(uint32_t count, int sm_count)
{
    static const int env = getenv("WVB_SPREAD") ? atoi(getenv("WVB_SPREAD")) : -1; // experiments: force an exponent
    if (env >= 0) return (uint32_t)std::min(env, 5);
    return count <= (uint32_t)sm_count * 4u ? 5u : 0u;
}

// shared-memory size classes (words per thread); a launch uses the smallest class that fits its blocks
const int kSmemClasses[] = {8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 128};
constexpr int SMEM_CLASS_MAX = 448, SMEM_CLASS_SMALL_CTA = 128; // classes above 128 words: steps of 8, one-warp CTAs
int smem_class(int words)
{
    for (int c : kSmemClasses)
        if (words <= c) return c;
    const int c = (words + 7) & ~7;
    return c <= SMEM_CLASS_MAX ? c : -1;
}

// Sort blocks so that warps are homogeneous; emit one launch per (variant, smem class).
void make_plan(const wvb_block_desc *descs, size_t n, int fmt, std::vector<uint32_t> &order, std::vector<Launch> &launches)
{
    order.resize(n);
    std::iota(order.begin(), order.end(), 0u);
    std::vector<uint64_t> key(n);
    bool any_checksum = false;
    for (size_t i = 0; i < n; i++) {
        const wvb_block_desc &d = descs[i];
        any_checksum |= (d.bflags & WVB_BF_BLOCK_CHECKSUM) != 0;
        int v = wvb::variant_of(d);
        // 16-bit interleaved stereo PCM gets its own launches: one aligned word store per frame, no byte-packing state
        if ((v & ~wvb::V_ANY_FIXED) == wvb::V_STEREO && wvb::block_is_fast16(d, fmt == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : fmt)) v |= wvb::V_F16;
        // (the staged-output kernels keep a 16-slot ring in the column; the in-register ones reuse dead state slots for it)
        const bool ring_in_class = WVB_STAGE_OUTPUT && (v & wvb::V_F16) && !(v & wvb::V_ANY_FIXED);
        int cls = v == wvb::V_DSD ? wvb::dsd_mode_class(d) : smem_class(std::max<int>(d.smem_words + (ring_in_class ? wvb::STAGE_RING_SLOTS : 0), (WVB_STAGE_OUTPUT && (v & wvb::V_F16)) ? wvb::STAGE_RING_SLOTS : 0));
        uint64_t clsbits = (uint64_t)(cls < 0 ? 1023 : cls) & 1023;
        // variant (10 bits) | class (10) | term signature (16) | inverted length (28; longer blocks saturate, they only lose
        // their order), so that long blocks start first
        const uint64_t inv_len = 0xfffffffu - (d.block_samples < 0xfffffffu ? d.block_samples : 0xfffffffu);
        key[i] = ((uint64_t)v << 54) | (clsbits << 44) | ((uint64_t)(d.terms_sig & 0xffff) << 28) | inv_len;
    }
    // Ties keep the table's order, i.e. neighbouring blocks of the same files share a warp: their compressed streams and
    // their outputs lie next to each other in the slabs.  (Measured and rejected in round 2: ordering ties by compressed
    // size so that silent stretches -- zero-run mode, which 64 % of the warp iterations execute for the sake of one lane --
    // get warps of their own.  The bench launch went from 77 to 150 ms: every lane of a warp then streams from a different
    // 2 MB page of the slabs and the TLBs thrash.)
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
    launches.clear();
    size_t i = 0;
    while (i < n) {
        uint64_t k = key[order[i]] >> 44;
        size_t j = i;
        while (j < n && (key[order[j]] >> 44) == k) j++;
        Launch L;
        L.variant = (int)(k >> 10);
        L.cls = (int)(k & 1023);
        L.first = (uint32_t)i;
        L.count = (uint32_t)(j - i);
        launches.push_back(L);
        i = j;
    }
    if (any_checksum && n) launches.push_back(Launch{wvb::V_CHECKSUM, 0, 0u, (uint32_t)n}); // after the decode launches: it ORs into their results
}

} // namespace

extern "C" {

int wvb_abi_version(void) { return WVB_ABI_VERSION; }
const char *wvb_last_error(void) { return g_last_error.c_str(); }

int wvb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int wvb_batch_create(int device, wvb_batch **out)
{
    if (!out) return WVB_E_ARG;
    *out = nullptr;
    int n = wvb_device_count();
    if (n <= 0) return set_error(WVB_E_NO_DEVICE, "no CUDA device: libwvb has no CPU decode path");
    if (device < 0 || device >= n) return set_error(WVB_E_ARG, "bad device ordinal");
    CUDA_TRY(cudaSetDevice(device));
    wvb_batch *b = new wvb_batch();
    b->device = device;
    CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    for (auto &e : b->ev) CUDA_TRY(cudaEventCreate(&e));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    b->sm_count = prop.multiProcessorCount;
    b->smem_optin = prop.sharedMemPerBlockOptin;
    *out = b;
    return WVB_OK;
}

void wvb_batch_destroy(wvb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->d_md5_ranges); cudaFree(b->d_md5_out);
    cudaFree(b->d_in); cudaFree(b->d_out); cudaFree(b->d_descs); cudaFree(b->d_order); cudaFree(b->d_results);
    cudaFree(b->d_scratch); cudaFree(b->d_scratch_meta);
    if (b->h_order) cudaFreeHost(b->h_order);
    if (b->h_descs) cudaFreeHost(b->h_descs);
    if (b->h_results) cudaFreeHost(b->h_results);
    for (auto &t : b->trace) { cudaEventDestroy(t.up); cudaEventDestroy(t.dec); cudaEventDestroy(t.down); }
    if (b->trace_t0) cudaEventDestroy(b->trace_t0);
    for (auto &e : b->ev) if (e) cudaEventDestroy(e);
    for (auto &e : b->seg_ev) if (e) cudaEventDestroy(e);
    for (auto &st : b->seg_streams) if (st) cudaStreamDestroy(st);
    if (b->s_in) cudaStreamDestroy(b->s_in);
    if (b->s_out) cudaStreamDestroy(b->s_out);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

void *wvb_batch_stream(wvb_batch *b) { return b ? (void *)b->stream : nullptr; }

int wvb_batch_wait(wvb_batch *b)
{
    if (!b) return WVB_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (b->pending_copy && b->pending_results) {
        memcpy(b->pending_results, b->h_results, b->pending_n * sizeof(wvb_block_result));
        b->pending_copy = false;
    }
    if (!b->trace.empty() && b->trace_t0) {
        fprintf(stderr, "[wvb trace] pipelined decode, %zu segments, host queueing %.2f ms; times in ms since the call began (device clock)\n",
                b->trace.size(), b->trace_host_total_ms);
        for (size_t k = 0; k < b->trace.size(); k++) {
            float up = 0, dec = 0, down = 0;
            cudaEventElapsedTime(&up, b->trace_t0, b->trace[k].up);
            cudaEventElapsedTime(&dec, b->trace_t0, b->trace[k].dec);
            cudaEventElapsedTime(&down, b->trace_t0, b->trace[k].down);
            fprintf(stderr, "[wvb trace] seg %2zu: %7zu blocks  in %7.1f MB  out %7.1f MB  queued@%7.2f  uploaded@%7.2f  decoded@%7.2f  downloaded@%7.2f\n", k,
                    b->trace[k].blocks, b->trace[k].in_bytes / 1e6, b->trace[k].out_bytes / 1e6, b->trace[k].host_ms, up, dec, down);
        }
        for (auto &t : b->trace) { cudaEventDestroy(t.up); cudaEventDestroy(t.dec); cudaEventDestroy(t.down); }
        b->trace.clear();
    }
    return WVB_OK;
}

int wvb_batch_timing(wvb_batch *b, float *kernel_ms, float *h2d_ms, float *d2h_ms, int *launches)
{
    if (!b || !b->timed) return WVB_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    float v = 0;
    if (h2d_ms) { CUDA_TRY(cudaEventElapsedTime(&v, b->ev[0], b->ev[1])); *h2d_ms = v; }
    if (kernel_ms) { CUDA_TRY(cudaEventElapsedTime(&v, b->ev[1], b->ev[2])); *kernel_ms = v; }
    if (d2h_ms) { CUDA_TRY(cudaEventElapsedTime(&v, b->ev[2], b->ev[3])); *d2h_ms = v; }
    if (launches) *launches = b->launches;
    return WVB_OK;
}

int wvb_batch_md5(wvb_batch *b, const void *device_out, size_t out_bytes, const uint64_t *offsets, const uint64_t *lengths, size_t n, uint8_t *digests)
{
    if (!b || !offsets || !lengths || !digests) return WVB_E_ARG;
    if (n == 0) return WVB_OK;
    if (n > 0xfffffff0ull) return WVB_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    const uint8_t *src = (const uint8_t *)device_out;
    if (!src) { // the batch's own copy of the last host-buffer decode
        src = b->d_out;
        if (!src || out_bytes > b->d_out_cap) return set_error(WVB_E_ARG, "wvb_batch_md5: no decoded output of that size is resident in the batch");
    }
    for (size_t i = 0; i < n; i++)
        if (offsets[i] > out_bytes || lengths[i] > out_bytes - offsets[i]) return set_error(WVB_E_ARG, "wvb_batch_md5: range outside the output slab");
    int rc;
    if ((rc = ensure(b->d_md5_ranges, b->d_md5_ranges_cap, 2 * n)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_md5_out, b->d_md5_out_cap, 16 * n)) != WVB_OK) return rc;
    cudaStream_t s = b->stream;
    CUDA_TRY(cudaMemcpyAsync(b->d_md5_ranges, offsets, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->d_md5_ranges + n, lengths, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    wvb::k_md5_ranges<<<(unsigned)((n + wvb::MD5_THREADS - 1) / wvb::MD5_THREADS), wvb::MD5_THREADS, 0, s>>>(src, b->d_md5_ranges, b->d_md5_ranges + n, (uint32_t)n,
                                                                                                            b->d_md5_out);
    CUDA_TRY(cudaGetLastError());
    b->launches++;
    CUDA_TRY(cudaMemcpyAsync(digests, b->d_md5_out, 16 * n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return WVB_OK;
}

static int validate_table(const wvb_block_desc *descs, size_t nblocks, size_t in_bytes, size_t out_bytes, int out_format)
{
    // Every descriptor must stay inside the slabs and inside its own frame: the table is the caller's word (wvb_index makes
    // well-formed ones, but a table can be hand-made or rebased wrongly) and the kernels trust it.  Comparisons are written
    // so that nothing wraps in 64 bits.
    const int ofmt = out_format == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : out_format;
    for (size_t i = 0; i < nblocks; i++) {
        const wvb_block_desc &d = descs[i];
        if (d.in_bytes > in_bytes || d.in_offset > in_bytes - d.in_bytes) return set_error(WVB_E_ARG, "descriptor input range outside the slab");
        if ((unsigned)d.out_ch_offset + d.out_channels > d.out_stride || d.out_channels == 0)
            return set_error(WVB_E_ARG, "descriptor channel slots outside its output frame");
        if (ofmt == WVB_OUT_PCM && (d.out_bps < 1 || d.out_bps > 4)) return set_error(WVB_E_ARG, "descriptor bytes per sample not in 1..4");
        const uint64_t fb = wvb_frame_bytes(&d, ofmt);
        // bytes written before out_offset: the zero-filled gap, and for DSD the mute fill, which starts at the caller chunk
        // that contains the end of the block and may therefore reach back into the previous block's output (DsdUtils.cs:99-117)
        uint64_t back = (uint64_t)d.gap_before * fb;
        if ((d.flags & 0x80000000u) && d.chunk_first != 0 && d.chunk_first < d.chunk_samples)
            back = std::max<uint64_t>(back, (uint64_t)(d.chunk_samples - d.chunk_first) * fb);
        const uint64_t fwd = (uint64_t)d.block_samples * fb; // < 2^32 * 1020
        if (d.out_offset > out_bytes || fwd > out_bytes - d.out_offset) return set_error(WVB_E_ARG, "descriptor output range outside the slab");
        if (back > d.out_offset) return set_error(WVB_E_ARG, "descriptor gap / mute fill starts before the slab");
        if (d.sub_len[WVB_SUB_TERMS] > 16) return set_error(WVB_E_ARG, "more than 16 decorrelation terms");
        for (int k = 0; k < WVB_SUB_COUNT; k++)
            if (d.sub_off[k] && ((uint64_t)d.sub_off[k] + d.sub_len[k] > d.in_bytes)) return set_error(WVB_E_ARG, "sub-block outside its block");
    }
    return WVB_OK;
}

static int upload_table(wvb_batch *b, const wvb_block_desc *descs, size_t nblocks, int fmt)
{
    int rc;
    make_plan(descs, nblocks, fmt, b->order, b->plan);
    if ((rc = ensure(b->d_descs, b->d_descs_cap, nblocks + 1)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_order, b->d_order_cap, nblocks + 1)) != WVB_OK) return rc;
    if (nblocks) {
        CUDA_TRY(cudaMemcpyAsync(b->d_descs, descs, nblocks * sizeof(wvb_block_desc), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->d_order, b->order.data(), nblocks * sizeof(uint32_t), cudaMemcpyHostToDevice, b->stream));
    }
    return WVB_OK;
}

// DSD fast-mode tables live in a scratch buffer sized once per table, BEFORE any launch (growing it later would free memory
// that kernels of an earlier segment still use)
static int ensure_dsd_scratch(wvb_batch *b, const std::vector<Launch> &plan)
{
    size_t slots = 0;
    for (const Launch &L : plan)
        if (L.variant == wvb::V_DSD && L.cls >= 16) slots = std::max(slots, (size_t)L.first + L.count);
    if (!slots) return WVB_OK;
    int rc;
    if (slots * wvb::DSD_FAST_TABLE_STRIDE > b->d_scratch_cap || slots * sizeof(wvb::DsdFastMeta) > b->d_scratch_meta_cap)
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    if ((rc = ensure(b->d_scratch, b->d_scratch_cap, slots * wvb::DSD_FAST_TABLE_STRIDE)) != WVB_OK) return rc;
    return ensure(b->d_scratch_meta, b->d_scratch_meta_cap, slots * sizeof(wvb::DsdFastMeta));
}

static int launch_plan(wvb_batch *b, const std::vector<Launch> &plan, const uint8_t *din, uint8_t *dout, int fmt, wvb_block_result *dres, cudaStream_t s)
{
    int rc;
    for (const Launch &L : plan) {
        if (L.variant == wvb::V_CHECKSUM) {
            constexpr unsigned per = wvb::CHECKSUM_THREADS / 32;
            wvb::k_block_checksum<<<(L.count + per - 1) / per, wvb::CHECKSUM_THREADS, 0, s>>>(din, b->d_descs, b->d_order + L.first, L.count, dres);
            CUDA_TRY(cudaGetLastError());
            b->launches++;
            continue;
        }
        if (L.variant == wvb::V_DSD) {
            // fast mode: scratch slots are indexed by position in the order array, so concurrent launches never share one
            if (L.cls >= 16 && ((size_t)L.first + L.count) * wvb::DSD_FAST_TABLE_STRIDE > b->d_scratch_cap)
                return set_error(WVB_E_ARG, "internal: DSD scratch not sized before launch");
            if ((rc = wvb::launch_dsd(L.cls, din, b->d_descs, b->d_order + L.first, L.count, dout, fmt, dres, s, b->smem_optin, b->device,
                                      &b->launches, b->d_scratch, b->d_scratch_meta, L.first)) != WVB_OK)
                return set_error(rc, std::string("DSD launch failed (unsupported dsd mode or CUDA error): ") + cudaGetErrorString(cudaGetLastError()));
            continue;
        }
        pcm_kernel_t k = pcm_kernel(L.variant);
        int cta = CTA_THREADS;
        if (L.variant & wvb::V_FIXED_D) cta = CTA_FIXED_D;
        else if (L.cls > SMEM_CLASS_SMALL_CTA) {
            pcm_kernel_t ks = pcm_kernel_small(L.variant);
            if (ks) { k = ks; cta = CTA_SMALL; }
        }
        if (!k || L.cls <= 0 || L.cls > SMEM_CLASS_MAX) return set_error(WVB_E_ARG, "block needs more decorrelation state than any kernel class provides");
        size_t smem = (size_t)L.cls * cta * sizeof(int) + ((WVB_STAGE_OUTPUT && (L.variant & wvb::V_F16)) ? (size_t)cta * 8 : 0);
        if (smem > b->smem_optin) return set_error(WVB_E_ARG, "shared memory class exceeds the device limit");
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // (left alone, the driver sizes the shared-memory carve-out for fewer CTAs than the state allows)
        if (L.cls > SMEM_CLASS_SMALL_CTA) CUDA_TRY(cudaFuncSetAttribute((const void *)k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        const uint32_t spread = pick_spread(L.count, b->sm_count);
        unsigned grid = (unsigned)((((uint64_t)L.count << spread) + cta - 1) / cta);
        k<<<grid, cta, smem, s>>>(din, b->d_descs, b->d_order + L.first, L.count, dout, fmt == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : fmt, dres, spread);
        CUDA_TRY(cudaGetLastError());
        b->launches++;
    }
    for (const Launch &L : plan) // 0x55 fill of muted DSD pieces, after every decode thread of this plan has finished
        if (L.variant == wvb::V_DSD && (rc = wvb::launch_dsd_mute_fix(b->d_descs, b->d_order + L.first, L.count, dout, fmt, dres, s, &b->launches)) != WVB_OK)
            return set_error(rc, "DSD mute pass failed");
    return WVB_OK;
}

// Host-buffer decode of a large table, pipelined: the table is cut into segments of consecutive descriptors; segment k's
// H2D copy, kernels and D2H copy run on three streams so that copies in both directions overlap the kernels of other
// segments (PCIe is the end-to-end bound: the PCM leaving the GPU is 2x the compressed bytes entering it).
// out_device: `out` is device memory (WVB_OUT_DEVICE): the kernels write into it directly and nothing is copied back, the
// upload still overlaps the decode segment by segment (the verify flow: host .wv bytes in, device PCM + MD5 out)
static int decode_pipelined_queue(wvb_batch *b, const uint8_t *in, size_t in_bytes, const wvb_block_desc *descs, size_t nblocks, uint8_t *out,
                                  size_t out_bytes, int fmt, wvb_block_result *results, bool out_device, bool *used);

static int decode_pipelined(wvb_batch *b, const uint8_t *in, size_t in_bytes, const wvb_block_desc *descs, size_t nblocks, uint8_t *out,
                            size_t out_bytes, int fmt, wvb_block_result *results, bool out_device, bool *used)
{
    const int rc = decode_pipelined_queue(b, in, in_bytes, descs, nblocks, out, out_bytes, fmt, results, out_device, used);
    if (rc != WVB_OK) {
        // Work may already be queued (uploads from `in`, kernels on the device slabs, downloads into `out`): the caller is
        // about to see an error and may free its buffers, so nothing may still be in flight when we return.
        const std::string msg = g_last_error;
        if (b->s_in) cudaStreamSynchronize(b->s_in);
        for (cudaStream_t st : b->seg_streams) cudaStreamSynchronize(st);
        if (b->s_out) cudaStreamSynchronize(b->s_out);
        cudaStreamSynchronize(b->stream);
        cudaGetLastError();
        b->plan.clear();
        b->order.clear();
        b->prepared = false;
        b->pending_copy = false;
        g_last_error = msg;
    }
    return rc;
}

static int decode_pipelined_queue(wvb_batch *b, const uint8_t *in, size_t in_bytes, const wvb_block_desc *descs, size_t nblocks, uint8_t *out,
                                  size_t out_bytes, int fmt, wvb_block_result *results, bool out_device, bool *used)
{
    *used = false;
    const int ofmt = fmt == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : fmt;
    size_t nseg = (in_bytes + out_bytes) / ((size_t)1536 << 20);
    if (nseg < 2 || nblocks < 4096) return WVB_OK;
    if (nseg > 16) nseg = 16;
    struct Seg { size_t first, count; uint64_t in_lo, in_hi, out_lo, out_hi; };
    std::vector<Seg> segs;
    // Segment sizes ramp up from a small first segment and down to a small last one (weights 1,2,3,4,...,4,3,2,1): the
    // first bytes can start their way back after the decode of a short segment (a block is a serial chain of ~30-50 ms
    // whatever the batch size, so that first decode cannot be hidden), and little is left to copy after the last kernel.
    std::vector<uint64_t> quota(nseg);
    {
        uint64_t wsum = 0;
        std::vector<uint64_t> wt(nseg);
        for (size_t k = 0; k < nseg; k++) { wt[k] = std::min<uint64_t>(std::min<uint64_t>(k + 1, nseg - k), 4); wsum += wt[k]; }
        const uint64_t total = in_bytes + out_bytes;
        for (size_t k = 0; k < nseg; k++) quota[k] = total / wsum * wt[k] + 1;
    }
    uint64_t acc = 0;
    Seg cur{0, 0, ~0ull, 0, ~0ull, 0};
    for (size_t i = 0; i < nblocks; i++) {
        const wvb_block_desc &d = descs[i];
        const uint64_t fb = wvb_frame_bytes(&d, ofmt);
        const uint64_t olo = d.out_offset - (uint64_t)d.gap_before * fb - (uint64_t)(wvb::variant_of(d) == wvb::V_DSD ? d.chunk_samples : 0) * fb;
        const uint64_t ohi = d.out_offset + (uint64_t)d.block_samples * fb;
        cur.in_lo = std::min<uint64_t>(cur.in_lo, d.in_offset); cur.in_hi = std::max<uint64_t>(cur.in_hi, d.in_offset + d.in_bytes);
        cur.out_lo = std::min<uint64_t>(cur.out_lo, std::min(olo, d.out_offset)); cur.out_hi = std::max(cur.out_hi, ohi);
        cur.count++;
        acc += d.in_bytes + (ohi - d.out_offset);
        if (acc >= quota[std::min(segs.size(), nseg - 1)] || i + 1 == nblocks) {
            segs.push_back(cur);
            cur = Seg{i + 1, 0, ~0ull, 0, ~0ull, 0};
            acc = 0;
        }
    }
    // the segments must tile the slabs: inputs nearly disjoint (files laid out in table order), outputs strictly disjoint
    uint64_t in_sum = 0;
    for (size_t k = 0; k < segs.size(); k++) {
        in_sum += segs[k].in_hi - segs[k].in_lo;
        if (k && segs[k].out_lo < segs[k - 1].out_hi) return WVB_OK;
        if (segs[k].out_hi > out_bytes || segs[k].in_hi > in_bytes) return WVB_OK;
    }
    if (in_sum > in_bytes + in_bytes / 4) return WVB_OK;

    int rc;
    if (!b->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
    if (!b->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
    while (b->seg_streams.size() < segs.size()) {
        cudaStream_t st;
        CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        b->seg_streams.push_back(st);
    }
    while (b->seg_ev.size() < 2 * segs.size() + 2) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->seg_ev.push_back(e);
    }
    if ((rc = ensure(b->d_in, b->d_in_cap, in_bytes + 64)) != WVB_OK) return rc;
    if (!out_device && (rc = ensure(b->d_out, b->d_out_cap, out_bytes + 64)) != WVB_OK) return rc;
    uint8_t *dout = out_device ? out : b->d_out;
    if ((rc = ensure(b->d_results, b->d_results_cap, nblocks + 1)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_descs, b->d_descs_cap, nblocks + 1)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_order, b->d_order_cap, nblocks + 1)) != WVB_OK) return rc;
    if ((rc = ensure_pinned(b->h_order, b->h_order_cap, nblocks + 1)) != WVB_OK) return rc;
    cudaStream_t s = b->stream;
    static const bool tracing = getenv("WVB_TRACE") && atoi(getenv("WVB_TRACE"));
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    for (auto &t : b->trace) { cudaEventDestroy(t.up); cudaEventDestroy(t.dec); cudaEventDestroy(t.down); }
    b->trace.clear();
    if (tracing) {
        if (!b->trace_t0) CUDA_TRY(cudaEventCreate(&b->trace_t0));
        CUDA_TRY(cudaEventRecord(b->trace_t0, s));
    }

    // Order of the host work matters here.  The table goes up first; then every segment is planned (a sort), its slice of
    // the launch order and its slab bytes are queued on the upload stream, and its kernels are launched -- so the first
    // launch waits for the planning of the (small) first segment only, and the planning of segment k+1 overlaps the upload
    // and decode of segment k.  (Small copies must not be queued behind the whole slab: the copy engine is first in, first out.)
    b->order.resize(nblocks);
    b->plan.clear();
    b->prepared = false;
    CUDA_TRY(cudaEventRecord(b->ev[0], s));
    CUDA_TRY(cudaMemcpyAsync(b->d_descs, descs, nblocks * sizeof(wvb_block_desc), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaEventRecord(b->seg_ev[2 * segs.size()], s));
    CUDA_TRY(cudaStreamWaitEvent(b->s_in, b->seg_ev[2 * segs.size()], 0));
    CUDA_TRY(cudaEventRecord(b->ev[1], s));

    std::vector<std::vector<Launch>> plans(segs.size());
    std::vector<uint32_t> tmp;
    auto plan_segment = [&](size_t k) {
        make_plan(descs + segs[k].first, segs[k].count, fmt, tmp, plans[k]);
        for (size_t j = 0; j < tmp.size(); j++) b->order[segs[k].first + j] = (uint32_t)(tmp[j] + segs[k].first);
        for (Launch &L : plans[k]) L.first += (uint32_t)segs[k].first;
    };
    // DSD fast mode keeps per-block tables in a scratch buffer that must be sized before the first launch (growing it later
    // would free memory a running kernel uses): such batches are planned up front
    bool needs_scratch = false;
    for (size_t i = 0; i < nblocks && !needs_scratch; i++) needs_scratch = (descs[i].flags & 0x80000000u) != 0;
    if (needs_scratch) {
        for (size_t k = 0; k < segs.size(); k++) plan_segment(k);
        for (const auto &pl : plans)
            if ((rc = ensure_dsd_scratch(b, pl)) != WVB_OK) return rc;
    }
    for (size_t k = 0; k < segs.size(); k++) {
        const Seg &g = segs[k];
        cudaStream_t ks = b->seg_streams[k];
        if (!needs_scratch) plan_segment(k);
        memcpy(b->h_order + g.first, b->order.data() + g.first, g.count * sizeof(uint32_t));
        CUDA_TRY(cudaMemcpyAsync(b->d_order + g.first, b->h_order + g.first, g.count * sizeof(uint32_t), cudaMemcpyHostToDevice, b->s_in));
        CUDA_TRY(cudaMemcpyAsync(b->d_in + g.in_lo, in + g.in_lo, g.in_hi - g.in_lo, cudaMemcpyHostToDevice, b->s_in));
        CUDA_TRY(cudaEventRecord(b->seg_ev[2 * k], b->s_in));
        wvb_batch::TraceSeg tr{nullptr, nullptr, nullptr, g.count, g.in_hi - g.in_lo, g.out_hi - g.out_lo, 0};
        if (tracing) {
            CUDA_TRY(cudaEventCreate(&tr.up)); CUDA_TRY(cudaEventCreate(&tr.dec)); CUDA_TRY(cudaEventCreate(&tr.down));
            CUDA_TRY(cudaEventRecord(tr.up, b->s_in));
        }
        CUDA_TRY(cudaStreamWaitEvent(ks, b->seg_ev[2 * k], 0)); // table (s_in waited for it), this segment's order slice and input
        if ((rc = launch_plan(b, plans[k], b->d_in, dout, fmt, b->d_results, ks)) != WVB_OK) return rc;
        CUDA_TRY(cudaEventRecord(b->seg_ev[2 * k + 1], ks));
        CUDA_TRY(cudaStreamWaitEvent(b->s_out, b->seg_ev[2 * k + 1], 0));
        CUDA_TRY(cudaStreamWaitEvent(s, b->seg_ev[2 * k + 1], 0));
        if (!out_device) CUDA_TRY(cudaMemcpyAsync(out + g.out_lo, dout + g.out_lo, g.out_hi - g.out_lo, cudaMemcpyDeviceToHost, b->s_out));
        if (tracing) {
            CUDA_TRY(cudaEventRecord(tr.dec, ks));
            CUDA_TRY(cudaEventRecord(tr.down, b->s_out));
            tr.host_ms = host_ms();
            b->trace.push_back(tr);
        }
    }
    b->trace_host_total_ms = host_ms();
    CUDA_TRY(cudaEventRecord(b->ev[2], s));
    CUDA_TRY(cudaEventRecord(b->seg_ev[2 * segs.size() + 1], b->s_out));
    CUDA_TRY(cudaStreamWaitEvent(s, b->seg_ev[2 * segs.size() + 1], 0));
    b->pending_copy = false;
    if (results) {
        if ((rc = ensure_pinned(b->h_results, b->h_results_cap, nblocks + 1)) != WVB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, nblocks * sizeof(wvb_block_result), cudaMemcpyDeviceToHost, s));
        b->pending_results = results;
        b->pending_n = nblocks;
        b->pending_copy = true;
    }
    CUDA_TRY(cudaEventRecord(b->ev[3], s));
    b->timed = true;
    *used = true;
    return WVB_OK;
}

int wvb_batch_prepare(wvb_batch *b, const wvb_block_desc *descs, size_t nblocks, int out_format)
{
    if (!b || !descs) return WVB_E_ARG;
    if (nblocks > 0xfffffff0ull) return WVB_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    b->prepared = false;
    int rc = validate_table(descs, nblocks, ~(size_t)0 >> 2, ~(size_t)0 >> 2, out_format);
    if (rc != WVB_OK) return rc;
    if ((rc = upload_table(b, descs, nblocks, out_format)) != WVB_OK) return rc;
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->prepared_in_extent = b->prepared_out_extent = 0;
    for (size_t i = 0; i < nblocks; i++) {
        const wvb_block_desc &d = descs[i];
        const uint64_t fb = wvb_frame_bytes(&d, out_format == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : out_format);
        b->prepared_in_extent = std::max<uint64_t>(b->prepared_in_extent, d.in_offset + d.in_bytes);
        b->prepared_out_extent = std::max<uint64_t>(b->prepared_out_extent, d.out_offset + (uint64_t)d.block_samples * fb);
    }
    b->prepared = true;
    b->prepared_n = nblocks;
    b->prepared_fmt = out_format;
    return WVB_OK;
}

int wvb_batch_decode(wvb_batch *b, const uint8_t *in, size_t in_bytes, const wvb_block_desc *descs, size_t nblocks, void *out,
                     size_t out_bytes, int out_format, uint32_t mem_flags, wvb_block_result *results)
{
    if (!b || !in || !out) return WVB_E_ARG;
    if (out_format != WVB_OUT_INT32 && out_format != WVB_OUT_PCM && out_format != WVB_OUT_DSD_RAW) return WVB_E_ARG;
    if (nblocks > 0xfffffff0ull) return WVB_E_ARG;
    if (!descs && !(b->prepared && b->prepared_n == nblocks && b->prepared_fmt == out_format))
        return set_error(WVB_E_ARG, "descs == NULL needs a matching wvb_batch_prepare");
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t s = b->stream;
    b->timed = false;
    b->launches = 0;
    int rc;
    if (descs) {
        b->prepared = false;
        if ((rc = validate_table(descs, nblocks, in_bytes, out_bytes, out_format)) != WVB_OK) return rc;
    } else if (b->prepared_in_extent > in_bytes || b->prepared_out_extent > out_bytes)
        return set_error(WVB_E_ARG, "slabs smaller than the prepared table needs");

    if (descs && !(mem_flags & (WVB_IN_DEVICE | WVB_RESULTS_DEVICE))) {
        bool used = false;
        if ((rc = decode_pipelined(b, in, in_bytes, descs, nblocks, (uint8_t *)out, out_bytes, out_format, results, (mem_flags & WVB_OUT_DEVICE) != 0,
                                   &used)) != WVB_OK)
            return rc;
        if (used) return (mem_flags & WVB_NO_SYNC) ? WVB_OK : wvb_batch_wait(b);
    }

    CUDA_TRY(cudaEventRecord(b->ev[0], s));
    const uint8_t *din = in;
    if (!(mem_flags & WVB_IN_DEVICE)) {
        if ((rc = ensure(b->d_in, b->d_in_cap, in_bytes + 64)) != WVB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->d_in, in, in_bytes, cudaMemcpyHostToDevice, s));
        din = b->d_in;
    }
    uint8_t *dout = (uint8_t *)out;
    if (!(mem_flags & WVB_OUT_DEVICE)) {
        if ((rc = ensure(b->d_out, b->d_out_cap, out_bytes + 64)) != WVB_OK) return rc;
        dout = b->d_out;
    }
    wvb_block_result *dres;
    if (mem_flags & WVB_RESULTS_DEVICE) {
        if (!results) return WVB_E_ARG;
        dres = results;
    } else {
        if ((rc = ensure(b->d_results, b->d_results_cap, nblocks + 1)) != WVB_OK) return rc;
        dres = b->d_results;
    }
    if (descs && (rc = upload_table(b, descs, nblocks, out_format)) != WVB_OK) return rc;
    CUDA_TRY(cudaEventRecord(b->ev[1], s));

    if ((rc = ensure_dsd_scratch(b, b->plan)) != WVB_OK) return rc;
    if ((rc = launch_plan(b, b->plan, din, dout, out_format, dres, s)) != WVB_OK) return rc;
    CUDA_TRY(cudaEventRecord(b->ev[2], s));

    if (!(mem_flags & WVB_OUT_DEVICE)) CUDA_TRY(cudaMemcpyAsync(out, dout, out_bytes, cudaMemcpyDeviceToHost, s));
    b->pending_copy = false;
    if (results && !(mem_flags & WVB_RESULTS_DEVICE)) {
        if ((rc = ensure_pinned(b->h_results, b->h_results_cap, nblocks + 1)) != WVB_OK) return rc;
        if (nblocks) CUDA_TRY(cudaMemcpyAsync(b->h_results, dres, nblocks * sizeof(wvb_block_result), cudaMemcpyDeviceToHost, s));
        b->pending_results = results;
        b->pending_n = nblocks;
        b->pending_copy = true;
    }
    CUDA_TRY(cudaEventRecord(b->ev[3], s));
    b->timed = true;
    if (!(mem_flags & WVB_NO_SYNC)) return wvb_batch_wait(b);
    return WVB_OK;
}

// Index + decode in one call, for host slabs: the index pass (host threads, file by file in slab order) runs WHILE the first
// segments are uploaded and decoded.  wvb_index_many + wvb_batch_decode put the whole index pass (20-25 ms for 10 000 files)
// in front of the first upload; here the first kernel launch waits for the files of the first, small segment only, and
// every later segment is launched as soon as the contiguous prefix of indexed files covers it.
int wvb_batch_decode_files(wvb_batch *b, const uint8_t *slab, size_t slab_bytes, const uint64_t *offsets, const uint64_t *sizes, size_t nfiles,
                           uint32_t open_flags, uint32_t chunk_samples, int out_format, int threads, wvb_file_info *infos,
                           wvb_block_desc *blocks, size_t cap, uint64_t *first, uint64_t *count, uint64_t *file_out_offset, size_t *nblocks,
                           uint64_t *out_bytes, void *out, size_t out_cap, uint32_t mem_flags, wvb_block_result *results)
{
    if (!b || !slab || !offsets || !sizes || !infos || !blocks || !first || !count || !file_out_offset || !out) return WVB_E_ARG;
    if (out_format != WVB_OUT_INT32 && out_format != WVB_OUT_PCM && out_format != WVB_OUT_DSD_RAW) return WVB_E_ARG;
    if (mem_flags & ~(uint32_t)(WVB_OUT_DEVICE | WVB_NO_SYNC)) return set_error(WVB_E_ARG, "wvb_batch_decode_files takes host input and host results");
    if (cap > 0xfffffff0ull) return WVB_E_ARG;
    for (size_t i = 0; i < nfiles; i++) {
        if (sizes[i] > slab_bytes || offsets[i] > slab_bytes - sizes[i]) return set_error(WVB_E_ARG, "file outside the slab");
        if (i && offsets[i] < offsets[i - 1] + sizes[i - 1]) return set_error(WVB_E_ARG, "files must be in slab order and disjoint");
    }
    CUDA_TRY(cudaSetDevice(b->device));
    const bool out_device = (mem_flags & WVB_OUT_DEVICE) != 0;
    const int ofmt = out_format == WVB_OUT_DSD_RAW ? WVB_OUT_PCM : out_format;
    cudaStream_t s = b->stream;
    b->timed = false;
    b->launches = 0;
    b->prepared = false;
    b->plan.clear();
    int rc;

    // segments by input bytes (the output size is not known yet), sizes ramping 1,2,3,4,...,4,3,2,1 as in decode_pipelined
    size_t nseg = (size_t)std::min<uint64_t>(16, std::max<uint64_t>(1, (uint64_t)slab_bytes * 3 / ((uint64_t)1536 << 20)));
    if (nfiles < 64) nseg = 1;
    std::vector<uint64_t> quota(nseg);
    {
        uint64_t wsum = 0, in_total = 0;
        for (size_t i = 0; i < nfiles; i++) in_total += sizes[i];
        std::vector<uint64_t> wt(nseg);
        for (size_t k = 0; k < nseg; k++) { wt[k] = std::min<uint64_t>(std::min<uint64_t>(k + 1, nseg - k), 4); wsum += wt[k]; }
        for (size_t k = 0; k < nseg; k++) quota[k] = in_total / wsum * wt[k] + 1;
    }
    if (!b->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking));
    if (!b->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
    while (b->seg_streams.size() < nseg + 1) {
        cudaStream_t st;
        CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        b->seg_streams.push_back(st);
    }
    while (b->seg_ev.size() < 2 * (nseg + 1) + 2) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->seg_ev.push_back(e);
    }
    if ((rc = ensure(b->d_in, b->d_in_cap, slab_bytes + 64)) != WVB_OK) return rc;
    if (!out_device && (rc = ensure(b->d_out, b->d_out_cap, out_cap + 64)) != WVB_OK) return rc;
    uint8_t *dout = out_device ? (uint8_t *)out : b->d_out;
    if ((rc = ensure(b->d_results, b->d_results_cap, cap + 1)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_descs, b->d_descs_cap, cap + 1)) != WVB_OK) return rc;
    if ((rc = ensure(b->d_order, b->d_order_cap, cap + 1)) != WVB_OK) return rc;
    if ((rc = ensure_pinned(b->h_order, b->h_order_cap, cap + 1)) != WVB_OK) return rc;
    if ((rc = ensure_pinned(b->h_descs, b->h_descs_cap, cap + 1)) != WVB_OK) return rc;
    static const bool tracing = getenv("WVB_TRACE") && atoi(getenv("WVB_TRACE"));
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    for (auto &t : b->trace) { cudaEventDestroy(t.up); cudaEventDestroy(t.dec); cudaEventDestroy(t.down); }
    b->trace.clear();
    if (tracing) {
        if (!b->trace_t0) CUDA_TRY(cudaEventCreate(&b->trace_t0));
        CUDA_TRY(cudaEventRecord(b->trace_t0, s));
    }
    CUDA_TRY(cudaEventRecord(b->ev[0], s));
    CUDA_TRY(cudaEventRecord(b->ev[1], s));
    CUDA_TRY(cudaEventRecord(b->seg_ev[2 * (nseg + 1)], s));
    CUDA_TRY(cudaStreamWaitEvent(b->s_in, b->seg_ev[2 * (nseg + 1)], 0)); // (work queued on the batch's stream before this call)

    // ---- index workers: files are handed out in order, each file's descriptors land in a block of its own ----
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(nfiles, 1));
    if (threads > 2) --threads; // the calling thread works too (table layout, planning, launches): keep to the caller's core budget
    std::vector<std::vector<wvb_block_desc>> per_file(nfiles);
    std::unique_ptr<std::atomic<uint8_t>[]> done(new std::atomic<uint8_t>[nfiles ? nfiles : 1]);
    for (size_t i = 0; i < nfiles; i++) done[i].store(0, std::memory_order_relaxed);
    std::atomic<size_t> next{0};
    std::atomic<bool> cancel{false};
    auto walk = [&]() {
        std::vector<wvb_block_desc> scratch(256);
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= nfiles) break;
            size_t n = 0;
            if (!cancel.load(std::memory_order_relaxed)) {
                for (;;) {
                    const int r = wvb_index(slab + offsets[i], (size_t)sizes[i], open_flags, chunk_samples, &infos[i], scratch.data(), scratch.size(), &n);
                    if (r != WVB_E_CAPACITY) break;
                    scratch.resize(n + n / 4 + 16);
                }
                per_file[i].assign(scratch.begin(), scratch.begin() + (ptrdiff_t)n);
            }
            count[i] = n;
            done[i].store(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> workers;
    for (int t = 0; t < threads; t++) workers.emplace_back(walk);
    auto finish_workers = [&]() { for (auto &t : workers) if (t.joinable()) t.join(); };
    auto drain = [&]() {
        if (b->s_in) cudaStreamSynchronize(b->s_in);
        for (cudaStream_t st : b->seg_streams) cudaStreamSynchronize(st);
        if (b->s_out) cudaStreamSynchronize(b->s_out);
        cudaStreamSynchronize(b->stream);
    };
    auto fail = [&](int code) {
        const std::string msg = g_last_error;
        cancel.store(true);
        finish_workers();
        drain();
        cudaGetLastError();
        b->pending_copy = false;
        g_last_error = msg;
        return code;
    };

    // ---- the main thread follows the contiguous prefix of indexed files, lays the table out and launches segments ----
    uint64_t total = 0, obytes = 0, seg_acc = 0;
    size_t seg_k = 0, seg_first_block = 0, seg_first_file = 0;
    bool too_small = false;
    std::vector<uint32_t> tmp_order;
    std::vector<Launch> plan;
    for (size_t i = 0; i < nfiles; i++) {
        while (!done[i].load(std::memory_order_acquire)) std::this_thread::yield();
        const size_t n = (size_t)count[i];
        first[i] = total;
        file_out_offset[i] = obytes;
        const uint32_t unit = out_format == WVB_OUT_INT32 ? 4u : (uint32_t)infos[i].bytes_per_sample;
        const uint32_t ch = (open_flags & WVB_OPEN_ALL_CHANNELS) ? (uint32_t)infos[i].num_channels
                                                                 : (uint32_t)(infos[i].reduced_channels > 0 ? infos[i].reduced_channels : infos[i].num_channels);
        const uint64_t fbytes = (uint64_t)infos[i].indexed_samples * unit * ch;
        if (total + n > cap || obytes + fbytes > out_cap) too_small = true; // keep counting: the caller gets the sizes needed
        if (!too_small && n) {
            memcpy(blocks + total, per_file[i].data(), n * sizeof(wvb_block_desc));
            wvb_rebase(blocks + total, n, offsets[i], obytes, out_format, (uint32_t)i);
        }
        std::vector<wvb_block_desc>().swap(per_file[i]);
        total += n;
        obytes = (obytes + fbytes + 15) & ~(uint64_t)15;
        seg_acc += sizes[i];
        const bool last = i + 1 == nfiles;
        if (too_small || !(seg_acc >= quota[std::min(seg_k, nseg - 1)] || last)) continue;
        // launch segment seg_k: files [seg_first_file, i], blocks [seg_first_block, total)
        const size_t bfirst = seg_first_block, bcount = (size_t)total - seg_first_block;
        const uint64_t in_lo = offsets[seg_first_file], in_hi = offsets[i] + sizes[i];
        const uint64_t out_lo = file_out_offset[seg_first_file], out_hi = std::min<uint64_t>(obytes, out_cap);
        seg_first_block = (size_t)total;
        seg_first_file = i + 1;
        seg_acc = 0;
        const size_t k = std::min(seg_k, nseg);
        seg_k++;
        if (!bcount) continue;
        if ((rc = validate_table(blocks + bfirst, bcount, slab_bytes, out_cap, out_format)) != WVB_OK) return fail(rc);
        make_plan(blocks + bfirst, bcount, out_format, tmp_order, plan);
        for (size_t j = 0; j < bcount; j++) b->h_order[bfirst + j] = (uint32_t)(tmp_order[j] + bfirst);
        for (Launch &L : plan) L.first += (uint32_t)bfirst;
        { // DSD fast-mode tables: the scratch may only grow while nothing runs
            size_t slots = 0;
            for (const Launch &L : plan)
                if (L.variant == wvb::V_DSD && L.cls >= 16) slots = std::max(slots, (size_t)L.first + L.count);
            if (slots * wvb::DSD_FAST_TABLE_STRIDE > b->d_scratch_cap || slots * sizeof(wvb::DsdFastMeta) > b->d_scratch_meta_cap) {
                drain();
                if ((rc = ensure(b->d_scratch, b->d_scratch_cap, std::max<size_t>(slots, cap) * wvb::DSD_FAST_TABLE_STRIDE)) != WVB_OK) return fail(rc);
                if ((rc = ensure(b->d_scratch_meta, b->d_scratch_meta_cap, std::max<size_t>(slots, cap) * sizeof(wvb::DsdFastMeta))) != WVB_OK) return fail(rc);
            }
        }
        memcpy(b->h_descs + bfirst, blocks + bfirst, bcount * sizeof(wvb_block_desc));
        cudaStream_t ks = b->seg_streams[k];
        auto cu = [&](cudaError_t e, const char *what) -> int {
            if (e == cudaSuccess) return WVB_OK;
            return set_error(WVB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        };
        if ((rc = cu(cudaMemcpyAsync(b->d_descs + bfirst, b->h_descs + bfirst, bcount * sizeof(wvb_block_desc), cudaMemcpyHostToDevice, b->s_in), "descs upload")) != WVB_OK) return fail(rc);
        if ((rc = cu(cudaMemcpyAsync(b->d_order + bfirst, b->h_order + bfirst, bcount * sizeof(uint32_t), cudaMemcpyHostToDevice, b->s_in), "order upload")) != WVB_OK) return fail(rc);
        if ((rc = cu(cudaMemcpyAsync(b->d_in + in_lo, slab + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, b->s_in), "slab upload")) != WVB_OK) return fail(rc);
        if ((rc = cu(cudaEventRecord(b->seg_ev[2 * k], b->s_in), "event")) != WVB_OK) return fail(rc);
        wvb_batch::TraceSeg tr{nullptr, nullptr, nullptr, bcount, in_hi - in_lo, out_hi - out_lo, 0};
        if (tracing) {
            cudaEventCreate(&tr.up); cudaEventCreate(&tr.dec); cudaEventCreate(&tr.down);
            cudaEventRecord(tr.up, b->s_in);
        }
        if ((rc = cu(cudaStreamWaitEvent(ks, b->seg_ev[2 * k], 0), "wait")) != WVB_OK) return fail(rc);
        if ((rc = launch_plan(b, plan, b->d_in, dout, out_format, b->d_results, ks)) != WVB_OK) return fail(rc);
        if ((rc = cu(cudaEventRecord(b->seg_ev[2 * k + 1], ks), "event")) != WVB_OK) return fail(rc);
        cudaStreamWaitEvent(b->s_out, b->seg_ev[2 * k + 1], 0);
        cudaStreamWaitEvent(s, b->seg_ev[2 * k + 1], 0);
        if (!out_device && out_hi > out_lo &&
            (rc = cu(cudaMemcpyAsync((uint8_t *)out + out_lo, dout + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, b->s_out), "download")) != WVB_OK)
            return fail(rc);
        if (tracing) {
            cudaEventRecord(tr.dec, ks);
            cudaEventRecord(tr.down, b->s_out);
            tr.host_ms = host_ms();
            b->trace.push_back(tr);
        }
        (void)ofmt;
    }
    finish_workers();
    b->trace_host_total_ms = host_ms();
    if (nblocks) *nblocks = (size_t)total;
    if (out_bytes) *out_bytes = obytes;
    if (too_small) {
        drain();
        b->pending_copy = false;
        return set_error(WVB_E_CAPACITY, "block table or output slab too small: *nblocks / *out_bytes hold the sizes needed");
    }
    CUDA_TRY(cudaEventRecord(b->ev[2], s));
    CUDA_TRY(cudaEventRecord(b->seg_ev[2 * (nseg + 1) + 1], b->s_out));
    CUDA_TRY(cudaStreamWaitEvent(s, b->seg_ev[2 * (nseg + 1) + 1], 0));
    b->pending_copy = false;
    if (results && total) {
        if ((rc = ensure_pinned(b->h_results, b->h_results_cap, (size_t)total + 1)) != WVB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(b->h_results, b->d_results, (size_t)total * sizeof(wvb_block_result), cudaMemcpyDeviceToHost, s));
        b->pending_results = results;
        b->pending_n = (size_t)total;
        b->pending_copy = true;
    }
    CUDA_TRY(cudaEventRecord(b->ev[3], s));
    b->timed = true;
    if (!(mem_flags & WVB_NO_SYNC)) return wvb_batch_wait(b);
    return WVB_OK;
}

int wvb_batch_dsd_to_dsf(wvb_batch *b, const void *device_src, size_t src_bytes, void *device_dst, size_t dst_bytes, const uint64_t *src_off,
                         const uint64_t *dst_off, const uint64_t *frames, const uint32_t *channels, size_t nfiles)
{
    if (!b || !device_src || !device_dst || !src_off || !dst_off || !frames || !channels) return WVB_E_ARG;
    if (!nfiles) return WVB_OK;
    if (nfiles > 0xffffffu) return WVB_E_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    std::vector<wvb::DsfJob> jobs(nfiles);
    uint64_t words = 0;
    for (size_t i = 0; i < nfiles; i++) {
        if (channels[i] == 0 || channels[i] > 255) return set_error(WVB_E_ARG, "wvb_batch_dsd_to_dsf: channel count not in 1..255");
        const uint64_t in_len = frames[i] * channels[i];
        const uint64_t out_len = (frames[i] + wvb::DSF_BLOCK - 1) / wvb::DSF_BLOCK * wvb::DSF_BLOCK * channels[i];
        if (src_off[i] > src_bytes || in_len > src_bytes - src_off[i] || dst_off[i] > dst_bytes || out_len > dst_bytes - dst_off[i] || (dst_off[i] & 3))
            return set_error(WVB_E_ARG, "wvb_batch_dsd_to_dsf: range outside the slabs (or destination not 4-byte aligned)");
        if (words > 0xffffffffull) return set_error(WVB_E_ARG, "wvb_batch_dsd_to_dsf: more than 16 GiB of DSF data in one call");
        jobs[i] = wvb::DsfJob{src_off[i], dst_off[i], frames[i], channels[i], (uint32_t)words};
        words += out_len / 4;
    }
    if (!words) return WVB_OK;
    int rc;
    if ((rc = ensure(b->d_md5_ranges, b->d_md5_ranges_cap, nfiles * sizeof(wvb::DsfJob) / sizeof(uint64_t) + 1)) != WVB_OK) return rc;
    cudaStream_t s = b->stream;
    CUDA_TRY(cudaMemcpyAsync(b->d_md5_ranges, jobs.data(), nfiles * sizeof(wvb::DsfJob), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s)); // (jobs is a local vector)
    wvb::k_dsd_to_dsf<<<(unsigned)((words + wvb::DSF_THREADS - 1) / wvb::DSF_THREADS), wvb::DSF_THREADS, 0, s>>>(
        (const uint8_t *)device_src, (uint8_t *)device_dst, (const wvb::DsfJob *)b->d_md5_ranges, (uint32_t)nfiles, words);
    CUDA_TRY(cudaGetLastError());
    b->launches++;
    CUDA_TRY(cudaStreamSynchronize(s));
    return WVB_OK;
}

void *wvb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void wvb_host_free(void *p) { if (p) cudaFreeHost(p); }

} // extern "C"
