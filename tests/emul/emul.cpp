// tests/emul/emul.cpp -- TEST HARNESS ONLY.  Compiles the device decode function of
// wavpackdecoder_b200/csrc/wvb_pcm.cuh for the host so its logic can be debugged and
// regression-tested without a GPU.  Never linked into libwvb.so; the product has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../wavpackdecoder_b200/csrc/wvb_pcm.cuh"
#include "../../wavpackdecoder_b200/csrc/wvb_plan.h"

namespace {
struct HostSM {
    std::vector<int> w;
    int &operator()(int i) { return w[(size_t)i]; }
};
}

extern "C" int emul_decode(const uint8_t *in, const wvb_block_desc *descs, size_t n, uint8_t *out, int out_format, wvb_block_result *results)
{
    for (size_t i = 0; i < n; i++) {
        const wvb_block_desc &D = descs[i];
        HostSM sm;
        sm.w.assign((size_t)D.smem_words + 64, 0);
        wvb_block_result r;
        memset(&r, 0, sizeof(r));
        switch (wvb::variant_of(D)) {
        case wvb::V_MONO: wvb::decode_block_pcm<false, false, false>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO: wvb::decode_block_pcm<true, false, false>(sm, in, D, out, out_format, &r); break;
        case wvb::V_MONO | wvb::V_GENFIX: wvb::decode_block_pcm<false, false, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO | wvb::V_GENFIX: wvb::decode_block_pcm<true, false, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID: wvb::decode_block_pcm<false, true, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID: wvb::decode_block_pcm<true, true, true>(sm, in, D, out, out_format, &r); break;
        default: r.rflags = WVB_RF_BAD_BLOCK; break;
        }
        if (results) results[i] = r;
    }
    return 0;
}
