// wvb_plan.h -- host-side launch planning shared by the CUDA batch code and the test emulation.
#pragma once
#include <stdint.h>

#include "../../include/wvb.h"

namespace wvb {

// kernel variants of the PCM path
enum { V_MONO = 0, V_STEREO = 1, V_GENFIX = 2, V_HYBRID = 4, V_DSD = 8, V_FIXED = 16, V_F16 = 32, V_FIXED_B = 64, V_FIXED_C = 128, V_FIXED_D = 256,
       V_COUNT = 512, V_ANY_FIXED = V_FIXED | V_FIXED_B | V_FIXED_C | V_FIXED_D,
       V_CHECKSUM = 200 /* not a decode variant: the block-checksum pass over a plan's blocks, queued after its decode launches */ };

// FNV-1a over the term list in DECODER order, as wvb_index computes wvb_block_desc.terms_sig
constexpr uint32_t terms_hash(const int *t, int n)
{
    uint32_t sig = 2166136261u;
    for (int k = 0; k < n; k++) sig = (sig ^ (uint32_t)(uint8_t)t[k]) * 16777619u;
    return sig;
}

// Term lists with an in-register kernel (decoder order = reverse of the file/encoder order).
//   stereo {18,18,2,3,-2}: WavPack's default-mode list (BASELINE configs[0]/[1]);  mono {18,18,2,3}: the same without the cross term
#define WVB_FIXED_STEREO_TERMS -2, 3, 2, 18, 18
#define WVB_FIXED_MONO_TERMS 3, 2, 18, 18
constexpr int kFixedStereo[] = {WVB_FIXED_STEREO_TERMS};
constexpr int kFixedMono[] = {WVB_FIXED_MONO_TERMS};
constexpr uint32_t kFixedStereoSig = terms_hash(kFixedStereo, 5);
constexpr uint32_t kFixedMonoSig = terms_hash(kFixedMono, 4);
//   stereo {18,18,2,17,3} (V_FIXED_B) and {18,17} (V_FIXED_C): what FFmpeg's encoder writes at its default and fastest
//   compression levels (tests/golden/ff_s16_stereo_c1.wv, ..._c0.wv); higher levels search a list per block and stay generic
#define WVB_FIXED_STEREO_B_TERMS 3, 17, 2, 18, 18
#define WVB_FIXED_STEREO_C_TERMS 17, 18
//   stereo {18,18,2,3,-2,18,2,4,7,5,3,6,8,-1,18,2} (V_FIXED_D): the 16-term list of libwavpack's "very high" mode, the list of
//   BASELINE configs[2]: 134 registers of weights and history, two 128-thread CTAs per SM -- and still twice as fast as the
//   shared-memory kernel, whose pass costs ~65 instructions against ~20 here
#define WVB_FIXED_STEREO_D_TERMS 2, 18, -1, 8, 6, 3, 5, 7, 4, 2, 18, -2, 3, 2, 18, 18
constexpr int kFixedStereoD[] = {WVB_FIXED_STEREO_D_TERMS};
constexpr uint32_t kFixedStereoDSig = terms_hash(kFixedStereoD, 16);
constexpr int kFixedStereoB[] = {WVB_FIXED_STEREO_B_TERMS};
constexpr int kFixedStereoC[] = {WVB_FIXED_STEREO_C_TERMS};
constexpr uint32_t kFixedStereoBSig = terms_hash(kFixedStereoB, 5);
constexpr uint32_t kFixedStereoCSig = terms_hash(kFixedStereoC, 2);

// 16-bit interleaved stereo PCM at a 4-byte aligned slab offset, the bench case: one aligned 32-bit store per frame, no
// byte packing state to carry (a caller-rebased table with odd offsets goes through the packed writer instead)
#ifdef __CUDACC__
__host__ __device__
#endif
inline bool block_is_fast16(const wvb_block_desc &D, int out_format)
{
    return out_format == WVB_OUT_PCM && D.out_bps == 2 && D.out_stride == 2 && D.out_ch_offset == 0 && (D.out_offset & 3u) == 0;
}

inline int variant_of(const wvb_block_desc &d)
{
    if (d.flags & 0x80000000u) return V_DSD;
    int v = (d.flags & (4u | 0x40000000u)) ? V_MONO : V_STEREO;
    if (d.flags & 8u) v |= V_HYBRID | V_GENFIX;
    else if ((d.flags & (0x80u | 0x100u)) || (d.bflags & WVB_BF_WVX_PRESENT)) v |= V_GENFIX;
    if (!(d.bflags & (WVB_BF_MUTE_ALL | WVB_BF_STALE_STATE))) {
        // use an in-register kernel when the block's term list is one we specialise (checked again on the device).  The two
        // stock lists also have float / int32 / hybrid instantiations; the other lists only plain lossless ones.
        const bool plain = !(v & V_GENFIX);
        const int base = v & V_STEREO;
        if (base == V_STEREO && d.sub_len[WVB_SUB_TERMS] == 5 && d.terms_sig == kFixedStereoSig) v |= V_FIXED;
        if (base == V_MONO && d.sub_len[WVB_SUB_TERMS] == 4 && d.terms_sig == kFixedMonoSig) v |= V_FIXED;
        if (plain && base == V_STEREO && d.sub_len[WVB_SUB_TERMS] == 5 && d.terms_sig == kFixedStereoBSig) v |= V_FIXED_B;
        if (plain && base == V_STEREO && d.sub_len[WVB_SUB_TERMS] == 2 && d.terms_sig == kFixedStereoCSig) v |= V_FIXED_C;
        if (plain && base == V_STEREO && d.sub_len[WVB_SUB_TERMS] == 16 && d.terms_sig == kFixedStereoDSig) v |= V_FIXED_D;
    }
    return v;
}

} // namespace wvb
