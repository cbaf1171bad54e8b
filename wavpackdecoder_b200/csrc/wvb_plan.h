// wvb_plan.h -- host-side launch planning shared by the CUDA batch code and the test emulation.
#pragma once
#include <stdint.h>

#include "../../include/wvb.h"

namespace wvb {

// kernel variants of the PCM path
enum { V_MONO = 0, V_STEREO = 1, V_GENFIX = 2, V_HYBRID = 4, V_DSD = 8, V_COUNT = 16 };

inline int variant_of(const wvb_block_desc &d)
{
    if (d.flags & 0x80000000u) return V_DSD;
    int v = (d.flags & (4u | 0x40000000u)) ? V_MONO : V_STEREO;
    if (d.flags & 8u) v |= V_HYBRID | V_GENFIX;
    else if ((d.flags & (0x80u | 0x100u)) || (d.bflags & WVB_BF_WVX_PRESENT)) v |= V_GENFIX;
    return v;
}

} // namespace wvb
