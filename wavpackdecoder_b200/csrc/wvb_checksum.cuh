// wvb_checksum.cuh -- WavPack 5 block checksum (ID_BLOCK_CHECKSUM, Defines.cs:83) verified on the device.
//
// The reference only notes the sub-block (MetadataUtils.cs:183).  Definition (WavPack 5 libwavpack, WavpackVerifySingleBlock):
// csum = 0xffffffff; csum = csum * 3 + w over the 16-bit little-endian words of the block from its 'wvpk' up to the
// checksum sub-block; stored as 4 bytes, or as the 2 bytes of csum ^ (csum >> 16).  The recurrence is affine mod 2^32, so a
// warp evaluates 512 bytes per step: 16-byte aligned vector loads, two words per 16x8-bit dot product, one warp reduction.
// The block bytes are already in HBM for the decode, so the check costs one more coalesced read of the compressed slab.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wvb.h"

namespace wvb {

constexpr int CHECKSUM_THREADS = 128;

static __device__ __forceinline__ uint32_t ck_pow3(uint32_t e)
{
    uint32_t r = 1, b = 3;
    while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
    return r;
}
// the recurrence over the 8 words of v, starting from 0: sum w_i * 3^(7-i)
static __device__ __forceinline__ uint32_t ck_fold16(const uint4 v)
{
    const uint32_t k = 0x0103u; // low half-word (first in memory) x 3 + high half-word x 1
    return ((__dp2a_lo(v.x, k, 0u) * 9u + __dp2a_lo(v.y, k, 0u)) * 9u + __dp2a_lo(v.z, k, 0u)) * 9u + __dp2a_lo(v.w, k, 0u);
}

static __global__ void __launch_bounds__(CHECKSUM_THREADS)
k_block_checksum(const uint8_t *__restrict__ in, const wvb_block_desc *__restrict__ descs, const uint32_t *__restrict__ order, uint32_t count,
                 wvb_block_result *__restrict__ results)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = (blockIdx.x * CHECKSUM_THREADS + threadIdx.x) >> 5;
    if (w >= count) return; // whole warp leaves together
    const uint32_t bi = order[w];
    const wvb_block_desc &D = descs[bi];
    if (!(D.bflags & WVB_BF_BLOCK_CHECKSUM)) return;
    const uint8_t *p = in + D.in_offset;
    const uint32_t nbytes = D.checksum_off & ~1u;
    bool bad = (uint64_t)nbytes + 6 > D.in_bytes + 2u; // sub-block header + at least 2 stored bytes inside the block
    uint32_t crc = 0xffffffffu;
    if (!bad) {
        // bytes before the first 16-byte aligned address go one word at a time (a block at an odd address: all of it)
        uint32_t head = (uint32_t)((16u - ((uintptr_t)p & 15u)) & 15u);
        if ((head & 1u) || head > nbytes) head = nbytes;
        for (uint32_t j = 0; j < head; j += 2) crc = crc * 3u + ((uint32_t)p[j] | ((uint32_t)p[j + 1] << 8)); // (warp-uniform)
        const uint32_t nchunks = (nbytes - head) >> 4;
        const uint4 *q = (const uint4 *)(p + head);
        const uint32_t lane_pow = ck_pow3(8u * (31u - lane)), step_pow = ck_pow3(256u);
        uint32_t c = 0;
        for (; c + 64 <= nchunks; c += 64) { // two steps per trip: both loads in flight
            const uint4 a0 = __ldg(q + c + lane), a1 = __ldg(q + c + 32 + lane);
            crc = crc * step_pow + __reduce_add_sync(0xffffffffu, ck_fold16(a0) * lane_pow);
            crc = crc * step_pow + __reduce_add_sync(0xffffffffu, ck_fold16(a1) * lane_pow);
        }
        for (; c + 32 <= nchunks; c += 32) crc = crc * step_pow + __reduce_add_sync(0xffffffffu, ck_fold16(__ldg(q + c + lane)) * lane_pow);
        const uint32_t rest = nchunks - c;
        if (rest) {
            const bool mine = lane < rest;
            const uint32_t term = mine ? ck_fold16(__ldg(q + c + lane)) * ck_pow3(8u * (rest - 1u - lane)) : 0u;
            crc = crc * ck_pow3(8u * rest) + __reduce_add_sync(0xffffffffu, term);
        }
        for (uint32_t j = head + (nchunks << 4); j < nbytes; j += 2) crc = crc * 3u + ((uint32_t)p[j] | ((uint32_t)p[j + 1] << 8));
        const uint8_t *st = p + nbytes + 2;
        const uint32_t stored_bytes = 2u * (uint32_t)p[nbytes + 1];
        if (stored_bytes == 4)
            bad = ((uint32_t)st[0] | ((uint32_t)st[1] << 8) | ((uint32_t)st[2] << 16) | ((uint32_t)st[3] << 24)) != crc;
        else {
            crc ^= crc >> 16;
            bad = stored_bytes != 2 || ((uint32_t)st[0] | ((uint32_t)st[1] << 8)) != (crc & 0xffffu);
        }
    }
    if (bad && lane == 0) atomicOr(&results[bi].rflags, WVB_RF_BLOCK_CHECKSUM);
}

} // namespace wvb
