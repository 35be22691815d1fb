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
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31u;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
