// host_params.h -- host-side derivation of the kernel parameters from the by-value configuration
// (include/ldpc_b200.h).  Shared by the C-ABI (ldpc_b200.cu) and by the CPU emulation of the kernel arithmetic
// (tools/emu/emu_decode.cpp, test infrastructure), so that both run the kernels with identical constants.
#pragma once
#include <algorithm>
#include <cstring>

#include "decode_kernels.cuh"
#include "ldpc_b200.h"

namespace ldpc {

inline int method_of(const ldpc_b200_config& c) { return (c.decode_method < 0 || c.decode_method > 5) ? 0 : c.decode_method; }

// FAID fast path (KIND_FAID_M / KIND_FAID_EF_M): every V2C LUT must be non-decreasing in |v| and identical for the
// four column-weight classes -- true for all LUT sets of the reference (CDecoder_FAID.cpp:12-165).
inline bool faid_luts_monotone(const ldpc_b200_config& c) {
    for (int it = 0; it < 6; ++it)
        for (int a = 0; a < 8; ++a) {
            for (int w = 1; w < 4; ++w)
                if (c.v2c_lut[it][w][a] != c.v2c_lut[it][0][a] || c.v2c_lut_ef[it][w][a] != c.v2c_lut_ef[it][0][a]) return false;
            if (a && (c.v2c_lut[it][0][a] < c.v2c_lut[it][0][a - 1] || c.v2c_lut_ef[it][0][a] < c.v2c_lut_ef[it][0][a - 1])) return false;
        }
    return true;
}

// which message-passing kernel a DecodeMethod runs (dispatch of CSimulate.cpp:136-164)
inline int kind_of(const ldpc_b200_config& c, bool allow_fast = true) {
    const int m = method_of(c);
    if (m == 0) return KIND_NMS;
    if (m == 1 || m == 3 || m == 4) return KIND_OMS;
    // erasure mode exists in the FAID + DTBF decoder only; in the hybrid one EF_ELIMINATION 2 just selects other thresholds
    // (CDecoder_FAID_2B1C.cpp:120-123)
    if (m == 2 && c.ef_elimination == 2) return KIND_FAID_ER;
    const bool fast = allow_fast && faid_luts_monotone(c);
    return c.ef_elimination ? (fast ? KIND_FAID_EF_M : KIND_FAID_EF) : (fast ? KIND_FAID_M : KIND_FAID);
}

inline int sat8(int x) { return x > 127 ? 127 : (x < -128 ? -128 : x); }

// cste as a function of the (clipped) minimum, as PRMT byte tables of 64 + cste -- the form the kernels use.
//   OMS_MODE 1 (CDecoder_OMS.cpp:386-432): "offset" lanes (norm) and "boost" lanes
//   OMS_MODE 0 (CDecoder_OMS.cpp:383-385): min(sat8(min - offset), 7) for every lane; negative for min < offset
inline void oms_tables(const ldpc_b200_config& c, uint32_t norm[2], uint32_t boost[2]) {
    const int F1 = (int8_t)c.factor_1, F2 = (int8_t)c.factor_2;
    uint8_t n[8], b[8];
    for (int m0 = 0; m0 < 8; ++m0) {
        int m = m0;
        if (c.oms_mode == 0) {
            m = std::min(sat8(m - (int8_t)c.oms_offset), 7);
            n[m0] = b[m0] = (uint8_t)(64 + m);
            continue;
        }
        if (m > F1) m = sat8(m - 1);
        if (m >= F2) m = sat8(m - 1);
        // negative results cannot occur for F1 >= 0; negative factors are rejected in create()
        n[m0] = (uint8_t)(64 + std::min(std::max(m, 0), 7));
        m = m0;
        if (m < F2) m = sat8(m + 1);
        if (m <= F1) m = sat8(m + 1);
        b[m0] = (uint8_t)(64 + std::min(m, 7));
    }
    auto pack = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    norm[0] = pack(n); norm[1] = pack(n + 4);
    boost[0] = pack(b); boost[1] = pack(b + 4);
}

// The single-instruction is-min select of the kernels needs cste_1 >= cste_2 for every reachable pair of minima
// (min2 >= min1).  True for every sane configuration (e.g. NMS Factor_1 <= Factor_2); otherwise the mask-select
// variant of the kernel is used.
inline bool select_is_monotone(int kind, const ldpc_b200_config& c, const uint32_t norm[2], const uint32_t boost[2]) {
    if (kind == KIND_NMS) {
        auto g = [](int m, int f) { unsigned p = (((unsigned)m & 0xFFu) * (unsigned)(f & 0xFFFF)) & 0xFFFFu; p >>= 5; return (int)(p < 7u ? p : 7u); };
        for (int x = 0; x <= 31; ++x)
            for (int y = x; y <= 31; ++y)
                if (g(y, c.factor_2) < g(x, c.factor_1)) return false;
        return true;
    }
    if (kind == KIND_OMS) {
        auto at = [](const uint32_t t[2], int i) { return (int)((t[i >> 2] >> (8 * (i & 3))) & 0xFF); };
        for (int i = 0; i < 7; ++i)
            if (at(norm, i + 1) < at(norm, i) || at(boost, i + 1) < at(boost, i)) return false;
        return true;
    }
    return true;  // FAID: cste = min(min, 7)
}


// V2C LUTs (CDecoder_FAID.cpp:12-165) as PRMT byte tables
inline void fill_lut_tables(const ldpc_b200_config& c, LutTables& lt) {
    auto pack = [](const int8_t* p) { return (uint32_t)(uint8_t)p[0] | ((uint32_t)(uint8_t)p[1] << 8) | ((uint32_t)(uint8_t)p[2] << 16) | ((uint32_t)(uint8_t)p[3] << 24); };
    for (int it = 0; it < 6; ++it)
        for (int w = 0; w < 4; ++w) {
            lt.lut[it][w][0] = pack(&c.v2c_lut[it][w][0]);
            lt.lut[it][w][1] = pack(&c.v2c_lut[it][w][4]);
            lt.lut_ef[it][w][0] = pack(&c.v2c_lut_ef[it][w][0]);
            lt.lut_ef[it][w][1] = pack(&c.v2c_lut_ef[it][w][4]);
        }
    // threshold tables of the fast path: thr[x] = largest y with LUT[y] == LUT[x], 127 when that is 7 (|v| is unbounded above)
    for (int it = 0; it < 6; ++it)
        for (int ef = 0; ef < 2; ++ef) {
            const int8_t* L = ef ? c.v2c_lut_ef[it][0] : c.v2c_lut[it][0];
            int8_t th[8];
            for (int x = 0; x < 8; ++x) {
                int y = x;
                while (y < 7 && L[y + 1] == L[x]) ++y;
                th[x] = (int8_t)(y == 7 ? 127 : y);
            }
            uint32_t* dst = ef ? lt.thr_ef[it] : lt.thr[it];
            dst[0] = pack(th);
            dst[1] = pack(th + 4);
        }
}

// Everything of DecParams that depends only on the configuration (buffers are filled in by the caller).
// Returns whether the single-instruction is-min select (MONO) may be used.
inline bool fill_dec_params(const ldpc_b200_config& c, int kind, int planes, DecParams& P) {
    memset(&P, 0, sizeof P);
    P.max_iter = c.max_iteration;
    P.planes = planes;
    P.hard2_thr = c.hard2_threshold;
    P.puncture_tail = c.puncture_tail;
    P.factor_1 = c.factor_1;
    P.factor_2 = c.factor_2;
    P.nms_fast = c.factor_1 >= 0 && c.factor_1 <= 2114 && c.factor_2 >= 0 && c.factor_2 <= 2114;  // 31 * 2114 < 65536
    oms_tables(c, P.oms_norm, P.oms_boost);
    P.oms_floor_err = (uint8_t)c.oms_floor_err_count;
    P.oms_floor_iter = c.oms_floor_iter_thresh;
    P.ef_floor_err = (int8_t)c.ef_floor_err_count;
    P.ef_floor_iter = c.ef_floor_iter_thresh;
    P.err_sat = (kind == KIND_OMS) ? 255 : 127;  // unsigned / signed saturating error_sum (CDecoder_OMS.cpp:113, CDecoder_FAID.cpp:294)
    fill_lut_tables(c, P.luts);
    return select_is_monotone(kind, c, P.oms_norm, P.oms_boost);
}

}  // namespace ldpc
