// cuda_emu_shim.h -- host (g++) stand-ins for the sm_100a intrinsics used by csrc/decode_kernels.cuh, so that the
// per-layer device functions can be compiled and executed on a CPU by tools/emu/emu_decode.cpp.
// TEST INFRASTRUCTURE ONLY (tests/test_kernel_emulation.py): it lets the exact kernel arithmetic be checked against
// the oracle without a GPU.  Never linked into the product library.
#pragma once
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static
#define __restrict__

static inline int16_t emu_lo(uint32_t x) { return (int16_t)(x & 0xFFFFu); }
static inline int16_t emu_hi(uint32_t x) { return (int16_t)(x >> 16); }
static inline uint32_t emu_mk(int lo, int hi) { return ((uint32_t)lo & 0xFFFFu) | (((uint32_t)hi & 0xFFFFu) << 16); }

static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 7u;
        r |= (uint32_t)((v >> (8 * sel)) & 0xFFu) << (8 * i);
    }
    return r;
}
// prmt.b32 generic mode: selector bit 3 replicates the msb of the selected byte
static inline uint32_t emu_prmt(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xFu;
        uint32_t byte = (uint32_t)((v >> (8 * (n & 7u))) & 0xFFu);
        if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t __vmins2(uint32_t a, uint32_t b) {
    return emu_mk(emu_lo(a) < emu_lo(b) ? emu_lo(a) : emu_lo(b), emu_hi(a) < emu_hi(b) ? emu_hi(a) : emu_hi(b));
}
static inline uint32_t __vmaxs2(uint32_t a, uint32_t b) {
    return emu_mk(emu_lo(a) > emu_lo(b) ? emu_lo(a) : emu_lo(b), emu_hi(a) > emu_hi(b) ? emu_hi(a) : emu_hi(b));
}
static inline uint32_t __vimin3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmins2(__vmins2(a, b), c); }
static inline uint32_t __vimax3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vmaxs2(a, b), c); }
static inline uint32_t __vadd2(uint32_t a, uint32_t b) { return emu_mk(emu_lo(a) + emu_lo(b), emu_hi(a) + emu_hi(b)); }
static inline uint32_t __vsub2(uint32_t a, uint32_t b) { return emu_mk(emu_lo(a) - emu_lo(b), emu_hi(a) - emu_hi(b)); }
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vadd2(a, b), c); }
static inline uint32_t __viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmins2(__vadd2(a, b), c); }
static inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vmins2(__vadd2(a, b), c), 0u); }
static inline uint32_t __viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) { return __vmaxs2(__vmaxs2(__vadd2(a, b), c), 0u); }
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const int x = (a >> (8 * i)) & 0xFF, y = (b >> (8 * i)) & 0xFF;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
}
// fp16 <-> float for the values the kernel uses (normal numbers and zero; results are exact small integers)
static inline float emu_h2f(uint16_t h) {
    const int sign = h >> 15, e = (h >> 10) & 31, m = h & 1023;
    float v;
    if (e == 0) v = (float)m * (1.0f / 16777216.0f);              // subnormal / zero
    else v = (1.0f + (float)m / 1024.0f) * (float)(1u << e) / 32768.0f;  // 2^(e-15)
    return sign ? -v : v;
}
static inline uint16_t emu_f2h(float f) {  // exact for integers |f| <= 2048 (all this code produces)
    if (f == 0.0f) return 0;
    const int sign = f < 0;
    float a = sign ? -f : f;
    int e = 15;
    while (a >= 2.0f) { a *= 0.5f; ++e; }
    while (a < 1.0f) { a *= 2.0f; --e; }
    const int m = (int)((a - 1.0f) * 1024.0f + 0.5f);
    return (uint16_t)((sign << 15) | (e << 10) | (m & 1023));
}
static inline uint32_t emu_h2_sub(uint32_t a, uint32_t b, bool sat) {
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        float v = emu_h2f((uint16_t)(a >> (16 * i))) - emu_h2f((uint16_t)(b >> (16 * i)));
        if (sat) v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
        r |= (uint32_t)emu_f2h(v) << (16 * i);
    }
    return r;
}
static inline uint32_t emu_h2_add(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i)
        r |= (uint32_t)emu_f2h(emu_h2f((uint16_t)(a >> (16 * i))) + emu_h2f((uint16_t)(b >> (16 * i)))) << (16 * i);
    return r;
}
static inline uint32_t emu_h2_fma(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r = 0;
    for (int i = 0; i < 2; ++i) {
        const float v = emu_h2f((uint16_t)(a >> (16 * i))) * emu_h2f((uint16_t)(b >> (16 * i))) + emu_h2f((uint16_t)(c >> (16 * i)));
        r |= (uint32_t)emu_f2h(v) << (16 * i);
    }
    return r;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
