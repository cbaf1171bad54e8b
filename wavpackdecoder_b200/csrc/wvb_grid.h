// wvb_grid.h -- the caller's call grid over one block, shared by the PCM decoder, the DSD decoders and the DSD mute pass.
//
// The reference decodes a block in caller-sized pieces, one unpack_samples / unpack_dsd_samples call each
// (WavPackUtils.cs:253-262), and three things depend on where the pieces start: the (short) casts of the decorrelation
// weights at the end of every pass call (UnpackUtils.cs:942-943,1152-1153,1239), the sample a mute starts from
// (UnpackUtils.cs:649-664, DsdUtils.cs:85-117) and the DSD 0x55 fill.  A descriptor carries the grid as
//   [0, skip_samples)            pieces of skip_chunk samples, the last one short: what seek() decodes and throws away
//                                in SAMPLE_BUFFER_SIZE / channels steps before the target (WavPackUtils.cs:573-578)
//   [skip_samples, ...)          a first piece of chunk_first samples (the rest of a call that began before the block;
//                                0 = a whole call), then pieces of chunk_samples
#pragma once
#include <stdint.h>

#include "../../include/wvb.h"

#if defined(__CUDACC__)
#define WVB_HD __host__ __device__ __forceinline__
#else
#define WVB_HD inline
#endif

namespace wvb {

// piece [ps, pe) of the block's first n samples that contains sample t (t < n)
WVB_HD void piece_of(const wvb_block_desc &D, uint32_t n, uint32_t t, uint32_t &ps, uint32_t &pe)
{
    const uint32_t skip = D.skip_samples < n ? D.skip_samples : n;
    if (t < skip) {
        const uint32_t k = D.skip_chunk ? D.skip_chunk : 0xffffffffu;
        ps = (t / k) * k;
        pe = (skip - ps) < k ? skip : ps + k;
        return;
    }
    const uint32_t chunk = D.chunk_samples ? D.chunk_samples : 0xffffffffu;
    const uint32_t m = n - skip, u = t - skip;
    uint32_t first = D.chunk_first < m ? D.chunk_first : m;
    if (first == 0) first = chunk < m ? chunk : m;
    if (u < first) {
        ps = skip;
        pe = skip + first;
        return;
    }
    const uint32_t s = first + ((u - first) / chunk) * chunk;
    ps = skip + s;
    pe = skip + ((m - s) < chunk ? m : s + chunk);
}

// samples of the call that produced piece [ps, ...) which were decoded BEFORE the block began (0 unless the block's first
// piece is the tail of a call that started in the previous block or in a gap): the DSD mute fill starts that far back
// (quirk C-11), and a failing first piece of that kind leaves stale caller data behind
WVB_HD uint32_t call_lookback(const wvb_block_desc &D, uint32_t ps)
{
    const uint32_t chunk = D.chunk_samples ? D.chunk_samples : 0xffffffffu;
    return (ps == 0 && D.skip_samples == 0 && D.chunk_first != 0 && D.chunk_first < chunk) ? chunk - D.chunk_first : 0;
}

} // namespace wvb
