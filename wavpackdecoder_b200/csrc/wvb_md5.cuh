// wvb_md5.cuh -- MD5 of byte ranges of the decoded output, on the device (integrity).
// WavPack stores the MD5 of the source audio data in ID_MD5_CHECKSUM (Defines.cs:77); the reference ignores it
// (MetadataUtils.cs:187-191 treats the optional id as unknown).  Checking it on the device means a verify-only pass
// never copies PCM back over PCIe.  One range (one file's PCM) per thread: MD5 is a serial chain per message, the
// parallelism is across files.  RFC 1321.
#pragma once
#include <stdint.h>

namespace wvb {

__device__ __constant__ uint32_t k_md5_t[64] = {
    0xd76aa478u, 0xe8c7b756u, 0x242070dbu, 0xc1bdceeeu, 0xf57c0fafu, 0x4787c62au, 0xa8304613u, 0xfd469501u, 0x698098d8u, 0x8b44f7afu, 0xffff5bb1u,
    0x895cd7beu, 0x6b901122u, 0xfd987193u, 0xa679438eu, 0x49b40821u, 0xf61e2562u, 0xc040b340u, 0x265e5a51u, 0xe9b6c7aau, 0xd62f105du, 0x02441453u,
    0xd8a1e681u, 0xe7d3fbc8u, 0x21e1cde6u, 0xc33707d6u, 0xf4d50d87u, 0x455a14edu, 0xa9e3e905u, 0xfcefa3f8u, 0x676f02d9u, 0x8d2a4c8au, 0xfffa3942u,
    0x8771f681u, 0x6d9d6122u, 0xfde5380cu, 0xa4beea44u, 0x4bdecfa9u, 0xf6bb4b60u, 0xbebfbc70u, 0x289b7ec6u, 0xeaa127fau, 0xd4ef3085u, 0x04881d05u,
    0xd9d4d039u, 0xe6db99e5u, 0x1fa27cf8u, 0xc4ac5665u, 0xf4292244u, 0x432aff97u, 0xab9423a7u, 0xfc93a039u, 0x655b59c3u, 0x8f0ccc92u, 0xffeff47du,
    0x85845dd1u, 0x6fa87e4fu, 0xfe2ce6e0u, 0xa3014314u, 0x4e0811a1u, 0xf7537e82u, 0xbd3af235u, 0x2ad7d2bbu, 0xeb86d391u};

__device__ __forceinline__ uint32_t md5_rotl(uint32_t x, int s) { return __funnelshift_l(x, x, s); }

// one 64-byte block, message words in m[16]
__device__ __forceinline__ void md5_block(uint32_t (&h)[4], const uint32_t (&m)[16])
{
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        uint32_t f;
        int g, s;
        if (i < 16) { f = (b & c) | (~b & d); g = i; s = (i & 3) == 0 ? 7 : (i & 3) == 1 ? 12 : (i & 3) == 2 ? 17 : 22; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; s = (i & 3) == 0 ? 5 : (i & 3) == 1 ? 9 : (i & 3) == 2 ? 14 : 20; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; s = (i & 3) == 0 ? 4 : (i & 3) == 1 ? 11 : (i & 3) == 2 ? 16 : 23; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; s = (i & 3) == 0 ? 6 : (i & 3) == 1 ? 10 : (i & 3) == 2 ? 15 : 21; }
        const uint32_t t = a + f + k_md5_t[i] + m[g];
        a = d; d = c; c = b;
        b = b + md5_rotl(t, s);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}

constexpr int MD5_THREADS = 64;

static __global__ void __launch_bounds__(MD5_THREADS)
k_md5_ranges(const uint8_t *__restrict__ data, const uint64_t *__restrict__ offsets, const uint64_t *__restrict__ lengths, uint32_t n,
             uint8_t *__restrict__ digests)
{
    const uint32_t i = blockIdx.x * MD5_THREADS + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = data + offsets[i];
    const uint64_t len = lengths[i];
    uint32_t h[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    uint32_t m[16];
    const bool aligned = ((uintptr_t)p & 3) == 0;
    uint64_t pos = 0;
    for (; pos + 64 <= len; pos += 64) {
        if (aligned) {
#pragma unroll
            for (int k = 0; k < 16; ++k) m[k] = *(const uint32_t *)(p + pos + 4 * k);
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint8_t *q = p + pos + 4 * k;
                m[k] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
            }
        }
        md5_block(h, m);
    }
    // tail: remaining bytes, 0x80, zero pad, 64-bit bit length -- one or two more blocks
    const int rem = (int)(len - pos);
    for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t w = 0;
            for (int j = 0; j < 4; ++j) {
                const int at = blk * 64 + 4 * k + j;
                uint32_t byte = 0;
                if (at < rem) byte = p[pos + at];
                else if (at == rem) byte = 0x80;
                w |= byte << (8 * j);
            }
            m[k] = w;
        }
        const bool last = blk == 1 || rem < 56;
        if (last) {
            const uint64_t bits = len * 8;
            m[14] = (uint32_t)bits;
            m[15] = (uint32_t)(bits >> 32);
        }
        md5_block(h, m);
        if (last) break;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint8_t *o = digests + 16ull * i + 4 * k;
        o[0] = (uint8_t)h[k]; o[1] = (uint8_t)(h[k] >> 8); o[2] = (uint8_t)(h[k] >> 16); o[3] = (uint8_t)(h[k] >> 24);
    }
}

} // namespace wvb
