// wvb_dsd.cuh -- DSD block decoders (DsdUtils.cs).  Placeholder: kernels land in the next commit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wvb.h"

namespace wvb {
inline int dsd_mode_class(const wvb_block_desc &) { return 0; }
inline int launch_dsd(int, const uint8_t *, const wvb_block_desc *, const uint32_t *, uint32_t, uint8_t *, int, wvb_block_result *,
                      cudaStream_t, size_t)
{
    return WVB_E_ARG;
}
} // namespace wvb
