// emu_decode.cpp -- CPU execution of the EXACT per-layer arithmetic of csrc/decode_kernels.cuh.
//
// TEST INFRASTRUCTURE ONLY (tests/test_kernel_emulation.py).  The device functions layer_0..layer_11 and the
// syndrome macros are compiled unchanged for the host through tools/emu/cuda_emu_shim.h; this file restates only
// the thin kernel body around them (LLR load, iteration loop, per-frame syndrome count, snapshot / group stop,
// hard decision), one "thread" after the other.  Threads of a layer touch disjoint APP words, so running them
// sequentially between the kernel's barriers is equivalent to the parallel execution.
// It exists so that changes to the integer tricks of the kernel can be checked bit-for-bit against the oracle
// in the GPU-less development container before any GPU time is spent.
#define LDPC_HOST_EMU 1
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "host_params.h"

using namespace ldpc;

namespace {

struct PairState {
    std::vector<uint32_t> app;           // [kN]
    std::vector<uint32_t> cv;            // [256][12][6]
    unsigned long long zmask[2] = {0, 0};
    std::vector<uint8_t> snap;           // [max_iter][2][kN] hard decisions at iteration starts with zero syndrome
    std::vector<uint8_t> final_hard;     // [2][kN]
};

inline void hard_of(const std::vector<uint32_t>& app, uint8_t* h0, uint8_t* h1, int bias) {
    for (int n = 0; n < kN; ++n) {
        h0[n] = (int)(int16_t)(app[n] & 0xFFFFu) > bias;
        h1[n] = (int)(int16_t)(app[n] >> 16) > bias;
    }
}

template <int KIND, bool MONO>
void run_pair(const DecParams& P, const int8_t* fix_group, int fg, PairState& st) {
    st.app.assign(pair_smem_words(KIND), 0);  // APP array (+ the syndrome words of KIND_FAID_ER behind the message area)
    st.cv.assign((size_t)256 * LDPC_MB * 6, 0x88888888u);
    st.snap.assign((size_t)P.max_iter * 2 * kN, 0);
    st.final_hard.assign((size_t)2 * kN, 0);
    uint32_t* app = st.app.data();
    constexpr int kB = bias_of(KIND);
    constexpr uint32_t hardk = hardk_of(KIND);
    (void)hardk;
    // load (reference two-region layout, CLDPC.cpp:234-272)
    for (int n = 0; n < kN; ++n) {
        int a, b;
        if (n < kK) {
            a = fix_group[(size_t)fg * kK + n];
            b = fix_group[(size_t)(fg + 1) * kK + n];
        } else {
            a = fix_group[(size_t)32 * kK + (size_t)fg * kM + (n - kK)];
            b = fix_group[(size_t)32 * kK + (size_t)(fg + 1) * kM + (n - kK)];
        }
        app[n] = pack_app(a, b, kB);
    }
    for (int n = kN - P.puncture_tail; n < kN; ++n) app[n] = pack_app(0, 0, kB);

    std::vector<uint32_t> chk0v(256, 0), chk1v(256, 0);
    for (int it = 1; it <= P.max_iter; ++it) {
        const int remaining = P.max_iter - it;
        IterCtx cxb;
        memset(&cxb, 0, sizeof cxb);
        if (KIND != KIND_NMS) {
            int e0 = 0, e1 = 0;
            for (int t = 0; t < 256; ++t) {
                const uint32_t rr = (uint32_t)t * 4u, pbase = 0;
                (void)rr; (void)pbase;
                uint32_t chk0 = 0, chk1 = 0;
                LDPC_FOR_EACH_LAYER(LDPC_SYN_LAYER)
                chk0v[t] = chk0;
                chk1v[t] = chk1;
                if (KIND == KIND_FAID_ER) app[unsat_word_offset(KIND) + t] = chk0 | (chk1 << 16);
                e0 += __popc(chk0);
                e1 += __popc(chk1);
            }
            const int err0 = e0 < P.err_sat ? e0 : P.err_sat;
            const int err1 = e1 < P.err_sat ? e1 : P.err_sat;
            if (err0 == 0) st.zmask[0] |= 1ull << (it - 1);
            if (err1 == 0) st.zmask[1] |= 1ull << (it - 1);
            if (err0 == 0 || err1 == 0)
                hard_of(st.app, &st.snap[((size_t)(it - 1) * 2) * kN], &st.snap[((size_t)(it - 1) * 2 + 1) * kN], kB);
            if (KIND == KIND_OMS) {
                cxb.lane_ok = expand2((unsigned)err0 < (unsigned)P.oms_floor_err, (unsigned)err1 < (unsigned)P.oms_floor_err);
                cxb.special_active = remaining <= P.oms_floor_iter;
            } else {
                cxb.lane_ok = expand2(err0 < P.ef_floor_err, err1 < P.ef_floor_err);
                cxb.special_active = kind_has_ef(KIND) && remaining <= P.ef_floor_iter;
            }
        }
        if (kind_is_faid(KIND)) {
            const int li = (it < 6 ? it : 6) - 1;
            for (int k = 0; k < 2; ++k) {
                cxb.thr[k] = P.luts.thr[li][k];
                cxb.thr_ef[k] = P.luts.thr_ef[li][k];
            }
            for (int w = 0; w < 4; ++w) {
                cxb.lut[w][0] = P.luts.lut[li][w][0];
                cxb.lut[w][1] = P.luts.lut[li][w][1];
                cxb.lut_ef[w][0] = P.luts.lut_ef[li][w][0];
                cxb.lut_ef[w][1] = P.luts.lut_ef[li][w][1];
            }
        }
#define EMU_RUN_LAYER(LY)                                                                              \
    for (int t = 0; t < 256; ++t) {                                                                    \
        IterCtx cx = cxb;                                                                              \
        cx.chk0 = chk0v[t];                                                                            \
        cx.chk1 = chk1v[t];                                                                            \
        uint32_t(&cvl)[6] = *reinterpret_cast<uint32_t(*)[6]>(&st.cv[((size_t)t * LDPC_MB + LY) * 6]); \
        uint32_t none[6];                                                                              \
        layer_##LY<KIND, MONO, false, false>(app, (uint32_t)t * 4u, 0u, cvl, nullptr, nullptr, none, cx, P);             \
    }
        LDPC_FOR_EACH_LAYER(EMU_RUN_LAYER)
#undef EMU_RUN_LAYER
    }
    hard_of(st.app, &st.final_hard[0], &st.final_hard[kN], kB);
}

}  // namespace

extern "C" {

// Runs the min-sum stage of one group of 32 frames exactly as decode_pair_kernel + the stop resolution of
// finalize_kernel would (no bit-flipping stage).  hard_out: int8[32][N] hard decisions (L > 0) at the group's stop
// point; its_executed: iterations executed by the group; conv_iter[32]: first iteration index with zero syndrome.
// Returns (kind << 4) | mono: the kernel kind that ran and whether the single-instruction (MONO) select was used.
int emu_decode_group(const ldpc_b200_config* cfg, int allow_fast, const int8_t* fix_group, int8_t* hard_out,
                     int32_t* its_executed, int32_t* conv_iter) {
    const int kind = kind_of(*cfg, allow_fast != 0);
    DecParams P;
    const bool mono = fill_dec_params(*cfg, kind, 1, P);
    std::vector<PairState> st(16);
    for (int p = 0; p < 16; ++p) {
        switch (kind) {
        case KIND_NMS:
            // with the fp16 select the NMS kernel's MONO flag carries nms_fast (decode_inst.cu)
            ((LDPC_FP16_SELECT ? (bool)P.nms_fast : mono)) ? run_pair<KIND_NMS, true>(P, fix_group, 2 * p, st[p]) : run_pair<KIND_NMS, false>(P, fix_group, 2 * p, st[p]);
            break;
        case KIND_OMS:
            mono ? run_pair<KIND_OMS, true>(P, fix_group, 2 * p, st[p]) : run_pair<KIND_OMS, false>(P, fix_group, 2 * p, st[p]);
            break;
        case KIND_FAID: run_pair<KIND_FAID, true>(P, fix_group, 2 * p, st[p]); break;
        case KIND_FAID_EF: run_pair<KIND_FAID_EF, true>(P, fix_group, 2 * p, st[p]); break;
        case KIND_FAID_M: run_pair<KIND_FAID_M, true>(P, fix_group, 2 * p, st[p]); break;
        case KIND_FAID_ER: run_pair<KIND_FAID_ER, true>(P, fix_group, 2 * p, st[p]); break;
        default: run_pair<KIND_FAID_EF_M, true>(P, fix_group, 2 * p, st[p]); break;
        }
    }
    // group stop: first iteration at whose start all 32 frames had a zero syndrome (CDecoder_OMS.cpp:325-327)
    int jstar = -1;
    if (kind != KIND_NMS)
        for (int j = 0; j < P.max_iter && jstar < 0; ++j) {
            int cnt = 0;
            for (int p = 0; p < 16; ++p) cnt += (int)((st[p].zmask[0] >> j) & 1) + (int)((st[p].zmask[1] >> j) & 1);
            if (cnt == 32) jstar = j;
        }
    for (int p = 0; p < 16; ++p)
        for (int f = 0; f < 2; ++f) {
            const uint8_t* src = jstar >= 0 ? &st[p].snap[((size_t)jstar * 2 + f) * kN] : &st[p].final_hard[(size_t)f * kN];
            memcpy(hard_out + (size_t)(2 * p + f) * kN, src, kN);
            if (conv_iter) {
                const unsigned long long z = st[p].zmask[f];
                int ci = z ? __builtin_ctzll(z) : -1;
                if (jstar >= 0 && ci > jstar) ci = -1;
                conv_iter[2 * p + f] = ci;
            }
        }
    if (its_executed) *its_executed = jstar >= 0 ? jstar : P.max_iter;
    return (mono ? 1 : 0) | (kind << 4);
}

}  // extern "C"
