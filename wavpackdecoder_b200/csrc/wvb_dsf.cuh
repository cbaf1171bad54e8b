// wvb_dsf.cuh -- re-layout of decoded DSD bytes for the Sony DSF container (container writer).
//
// The decoders produce what DSDIFF stores: one byte per channel per byte-time, channels interleaved, oldest bit in the MSB.
// DSF stores each channel in blocks of 4096 bytes (channel 0's block, channel 1's block, ..., then the next 4096
// byte-times), oldest bit in the LSB, the last block zero padded.  One thread moves four output bytes: a strided read of
// the interleaved stream, a bit reversal per byte, one aligned 32-bit store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wvb {

constexpr int DSF_THREADS = 256;
constexpr uint32_t DSF_BLOCK = 4096;

struct DsfJob { uint64_t src_off, dst_off, frames; uint32_t channels, first_word; }; // first_word: this file's first output word in the grid

static __global__ void __launch_bounds__(DSF_THREADS)
k_dsd_to_dsf(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const DsfJob *__restrict__ jobs, uint32_t njobs, uint64_t total_words)
{
    const uint64_t w = (uint64_t)blockIdx.x * DSF_THREADS + threadIdx.x;
    if (w >= total_words) return;
    uint32_t lo = 0, hi = njobs - 1; // the file this output word belongs to
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (jobs[mid].first_word <= w) lo = mid;
        else hi = mid - 1;
    }
    const DsfJob J = jobs[lo];
    const uint64_t byte = (w - J.first_word) * 4;           // offset inside the file's DSF data
    const uint64_t blk = byte / DSF_BLOCK;                  // 4096-byte block index: block group * channels + channel
    const uint32_t k = (uint32_t)(byte % DSF_BLOCK);
    const uint64_t group = blk / J.channels;
    const uint32_t c = (uint32_t)(blk % J.channels);
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t frame = group * DSF_BLOCK + k + i;
        const uint32_t v = frame < J.frames ? src[J.src_off + frame * J.channels + c] : 0u;
        out |= (__brev(v) >> 24) << (8 * i);
    }
    *(uint32_t *)(dst + J.dst_off + byte) = out;
}

} // namespace wvb
