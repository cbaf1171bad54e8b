// tests/emul/emul.cpp -- TEST HARNESS ONLY.  Compiles the device decode function of
// wavpackdecoder_b200/csrc/wvb_pcm.cuh for the host so its logic can be debugged and
// regression-tested without a GPU.  Never linked into libwvb.so; the product has no CPU path.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../wavpackdecoder_b200/csrc/wvb_dsd_core.cuh"
#include "../../wavpackdecoder_b200/csrc/wvb_pcm.cuh"
#include "../../wavpackdecoder_b200/csrc/wvb_plan.h"

namespace {
struct HostSM {
    std::vector<int> w;
    int cap_words = 0;
    int &operator()(int i) { return w.at((size_t)i); } // bounds-checked: an out-of-range slot index is a device fault
    int cap() const { return cap_words; }
};
}

extern "C" int emul_decode(const uint8_t *in, const wvb_block_desc *descs, size_t n, uint8_t *out, int out_format, wvb_block_result *results)
{
    for (size_t i = 0; i < n; i++) {
        const wvb_block_desc &D = descs[i];
        HostSM sm;
        sm.w.assign((size_t)D.smem_words + 1, 0);
        sm.cap_words = (int)D.smem_words;
        wvb_block_result r;
        memset(&r, 0, sizeof(r));
        switch (wvb::variant_of(D)) {
        case wvb::V_MONO: wvb::decode_block_pcm<false, false, false>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO:
            if (wvb::block_is_fast16(D, out_format)) wvb::decode_block_pcm<true, false, false, HostSM, wvb::GenericDecorr<true>, true>(sm, in, D, out, out_format, &r);
            else wvb::decode_block_pcm<true, false, false>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_MONO | wvb::V_FIXED:
            wvb::decode_block_pcm<false, false, false, HostSM, wvb::FixedDecorr<false, WVB_FIXED_MONO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_FIXED:
            if (wvb::block_is_fast16(D, out_format)) wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_TERMS>, true>(sm, in, D, out, out_format, &r);
            else wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_FIXED_B:
            if (wvb::block_is_fast16(D, out_format)) wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_B_TERMS>, true>(sm, in, D, out, out_format, &r);
            else wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_B_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_FIXED_C:
            if (wvb::block_is_fast16(D, out_format)) wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_C_TERMS>, true>(sm, in, D, out, out_format, &r);
            else wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_C_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_FIXED_D:
            if (wvb::block_is_fast16(D, out_format)) wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_D_TERMS>, true>(sm, in, D, out, out_format, &r);
            else wvb::decode_block_pcm<true, false, false, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_D_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_MONO | wvb::V_GENFIX | wvb::V_FIXED:
            wvb::decode_block_pcm<false, false, true, HostSM, wvb::FixedDecorr<false, WVB_FIXED_MONO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_FIXED:
            wvb::decode_block_pcm<true, false, true, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID | wvb::V_FIXED:
            wvb::decode_block_pcm<false, true, true, HostSM, wvb::FixedDecorr<false, WVB_FIXED_MONO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID | wvb::V_FIXED:
            wvb::decode_block_pcm<true, true, true, HostSM, wvb::FixedDecorr<true, WVB_FIXED_STEREO_TERMS>>(sm, in, D, out, out_format, &r);
            break;
        case wvb::V_MONO | wvb::V_GENFIX: wvb::decode_block_pcm<false, false, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO | wvb::V_GENFIX: wvb::decode_block_pcm<true, false, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_MONO | wvb::V_GENFIX | wvb::V_HYBRID: wvb::decode_block_pcm<false, true, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_STEREO | wvb::V_GENFIX | wvb::V_HYBRID: wvb::decode_block_pcm<true, true, true>(sm, in, D, out, out_format, &r); break;
        case wvb::V_DSD: {
            const int mode = wvb::dsd_key_mode(D.smem_words);
            if (mode == 0)
                wvb::dsd_decode_raw(in, D, out, out_format, &r);
            else if (mode == 3) {
                int pt0[256];
                wvb::dsd_init_ptable_host(pt0, wvb::dsd_key_rate(D.smem_words), 20);
                HostSM pt;
                pt.w.assign(256, 0);
                wvb::dsd_decode_high(pt, pt0, in, D, out, out_format, &r, true);
            } else if (mode == 1) {
                std::vector<uint16_t> summed(32 * 256);
                wvb::DsdFastTables T{summed.data()};
                const uint8_t *p = in + D.in_offset + D.sub_off[WVB_SUB_DSD];
                const uint32_t len = D.sub_len[WVB_SUB_DSD];
                int bins = 1;
                uint32_t at = wvb::dsd_fast_build(T, p, len, 0, 1, bins);
                if (!at) { r.rflags = WVB_RF_BAD_BLOCK; break; }
                wvb::dsd_fast_sums(T, bins, 0, 1, [](uint32_t) { return 0u; });
                wvb::DsdOut o;
                wvb::dsd_out_init(o, D, out, out_format);
                const bool mono = o.coded_ch == 1;
                const uint32_t total = D.block_samples * (uint32_t)o.coded_ch;
                int crc = -1;
                bool failed = false;
                uint32_t fail_at = total;
                wvb::dsd_fast_decode(T, bins, p, len, at, mono, total,
                    [](const uint16_t *row, int) { return (uint32_t)row[255]; },
                    [](const uint16_t *row, int, uint32_t index, uint32_t &below, uint32_t &cur) {
                        int c = 0;
                        for (int k = 0; k < 256; k++) c += row[k] <= index;
                        below = c > 0 ? row[c - 1] : 0u;
                        cur = row[c];
                        return c;
                    },
                    [&](uint32_t j, int code) { o.put(j, code); }, [](uint32_t) {}, crc, failed, fail_at);
                wvb::dsd_finish(D, &r, crc, failed, mono ? fail_at : fail_at >> 1, 0);
            } else
                r.rflags = WVB_RF_BAD_BLOCK;
            break;
        }
        default: r.rflags = WVB_RF_BAD_BLOCK; break;
        }
        if (results) results[i] = r;
    }
    // DSD mute pass (k_dsd_mute_fix): 0x55 fill of muted pieces
    for (size_t i = 0; i < n; i++) {
        const wvb_block_desc &D = descs[i];
        if (wvb::variant_of(D) != wvb::V_DSD || !results || !(results[i].rflags & WVB_RF_MUTED)) continue;
        const int unit = out_format == WVB_OUT_INT32 ? 4 : 1, add = out_format == WVB_OUT_PCM ? 128 : 0;
        const uint32_t fb = (uint32_t)unit * D.out_stride, nn = D.block_samples, chunk = D.chunk_samples ? D.chunk_samples : 0xffffffffu;
        uint32_t first_len = D.chunk_first < nn ? D.chunk_first : nn;
        if (first_len == 0) first_len = chunk < nn ? chunk : nn;
        uint32_t ps = results[i].mute_from;
        while (ps < nn) {
            const uint32_t pe = ps < first_len ? first_len : (ps + chunk < nn ? ps + chunk : nn);
            int64_t start = ps;
            if (ps == 0 && D.chunk_first != 0 && D.chunk_first < chunk) start = -(int64_t)(chunk - D.chunk_first);
            uint8_t *q = out + D.out_offset + start * (int64_t)fb;
            for (uint32_t k = 0; k < pe - ps; ++k, q += fb)
                for (int c = 0; c < D.out_stride; ++c) wvb::store_unit(q + c * unit, 0x55, unit, add);
            ps = pe;
        }
    }
    return 0;
}
