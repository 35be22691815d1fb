// gen_device.cuh -- device functions of the frame producer, shared by generate_kernel (frame_kernels.cuh) and by the
// fused producer->decoder loader of decode_pair_kernel (decode_kernels.cuh):
//   BeforeModulationInterleaver + Modulation   CModulate.cpp:95-152,216-264
//   CChannel::AWGNChannel                      CChannel.cpp:71-97   (Philox4x32-10 + Box-Muller instead of the 3-LCG stream)
//   Demodulation + AfterDeModulationDeInterleaver + float2LimitChar_4bit   CModulate.cpp:156-212,270-362; CLDPC.cpp:4524-4582
//
// Included from the middle of decode_kernels.cuh (after the code constants and GenCore, before the kernel).
#pragma once

namespace ldpc {

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so frame i always sees the same noise regardless of
// which GPU / stream / chunk processes it.
// ---------------------------------------------------------------------------------------------------------
struct Philox {
    static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
        const uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    __host__ __device__ static inline void gen(uint64_t seed, uint64_t subseq, uint64_t offset, uint32_t (&out)[4]) {
        uint32_t c[4] = {(uint32_t)offset, (uint32_t)(offset >> 32), (uint32_t)subseq, (uint32_t)(subseq >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += W0;
            k1 += W1;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

constexpr uint64_t kNoiseStream = 0;      // Philox offset space: symbol index
constexpr uint64_t kInfoStream = 1ull << 40;  // Philox offset space for info bits

// CModulate.cpp:4-6 (Gray maps)
static __constant__ float c_tab_qpsk[2] = {-0.707107f, 0.707107f};
static __constant__ float c_tab_16qam[4] = {-0.316228f, -0.948683f, 0.316228f, 0.948683f};
static __constant__ float c_tab_64qam[8] = {-0.462910f, -0.154303f, -0.771517f, -1.08012f, 0.462910f, 0.154303f, 0.771517f, 1.08012f};
static __constant__ float c_tab_256qam[16] = {-0.383482f, -0.536875f, -0.230089f, -0.076696f, -0.843661f, -0.690268f, -0.997054f, -1.150447f,
                                       0.383482f, 0.536875f, 0.230089f, 0.076696f, 0.843661f, 0.690268f, 0.997054f, 1.150447f};  // CModulate.cpp:7
__device__ __forceinline__ const float* gray_table(int mod) {
    return mod == 2 ? c_tab_qpsk : mod == 4 ? c_tab_16qam : mod == 6 ? c_tab_64qam : c_tab_256qam;
}
constexpr int kMaxMod = 8;

// float2LimitChar_4bit on one value.  _mm256_cvttps_epi32 yields INT_MIN for NaN / |x| >= 2^31, which the
// saturating packs and the clamp turn into -7 (CLDPC.cpp:4555-4573).
__device__ __forceinline__ int quant4(float x, float scale) {
    const float p = __fmul_rn(x, scale);
    if (!(p >= -2147483648.0f && p < 2147483648.0f)) return -7;
    const int t = __float2int_rz(p);
    return max(-7, min(7, t));
}

// float2LimitChar_{1,2,3,4,5,6}bit (CLDPC.cpp:4385-4770) on one value; see ldpc_b200_quantize_bits.
__device__ __forceinline__ int quant_bits(float x, float scale, int bits) {
    const float p = __fmul_rn(x, scale);
    int t = -2147483647 - 1;  // "integer indefinite"
    if (p >= -2147483648.0f && p < 2147483648.0f) t = bits == 6 ? __float2int_rn(p) : __float2int_rz(p);
    if (bits == 1) return t > 0 ? 31 : -31;
    const int lo = bits == 6 ? -31 : bits == 5 ? -16 : bits == 4 ? -7 : bits == 3 ? -4 : -2;
    const int hi = bits == 6 ? 31 : bits == 5 ? 15 : bits == 4 ? 7 : bits == 3 ? 3 : 1;
    return max(lo, min(hi, t));
}
// the quantiser selected by the configuration (GenCore.qbits; 4 = the reference's default path)
__device__ __forceinline__ int quant_cfg(float x, float scale, int qbits) {
    return qbits == 4 ? quant4(x, scale) : quant_bits(x, scale, qbits);
}

// position of transmitted-stream index `src` (within one frame) after AfterDeModulationDeInterleaver and the
// regrouping into the two-region layout; returns the byte offset inside the group's 32*N buffer.
__device__ __forceinline__ int deint_offset(int frame, int src, int I) {
    const int i = src / I, j = src - i * I;
    const int dst = j * (kN / I) + i;  // CModulate.cpp:161-172
    return dst < kK ? frame * kK + dst : 32 * kK + frame * kM + (dst - kK);  // :176-202
}

// max-log demapper without noise-variance scaling (CModulate.cpp:270-336); the subtractions are evaluated in
// double and rounded to float exactly like `fabs(float) - double_constant` in the reference.
__device__ __forceinline__ void demap_symbol(float re, float im, int mod, float (&llr)[kMaxMod]) {
    llr[0] = re;
    llr[1] = im;
    if (mod == 4) {
        llr[2] = (float)(fabs((double)re) - 0.6324555);
        llr[3] = (float)(fabs((double)im) - 0.6324555);
    } else if (mod == 6) {
        llr[2] = (float)(fabs((double)re) - 0.6172134);
        llr[3] = (float)(fabs((double)im) - 0.6172134);
        llr[4] = (float)(fabs((double)llr[2]) - 0.3086067);
        llr[5] = (float)(fabs((double)llr[3]) - 0.3086067);
    } else if (mod == 8) {  // CModulate.cpp:340-356
        llr[2] = (float)(fabs((double)re) - 0.613568);
        llr[3] = (float)(fabs((double)im) - 0.613568);
        llr[4] = (float)(fabs((double)llr[2]) - 0.306784);
        llr[5] = (float)(fabs((double)llr[3]) - 0.306784);
        llr[6] = (float)(fabs((double)llr[4]) - 0.153392);
        llr[7] = (float)(fabs((double)llr[5]) - 0.153392);
    }
}


// ---- transmitted symbols -----------------------------------------------------------------------------------------
// Gray-mapped symbols 2*sp and 2*sp+1 of one frame, general interleaver (CModulate.cpp:137-149, 243-262).
__device__ __forceinline__ void map_symbol_pair(const GenCore& G, int group, int frame, int sp, float (&re)[2], float (&im)[2]) {
    const int half = G.mod / 2;
    const float* tab = gray_table(G.mod);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int sf = 2 * sp + k;
        unsigned ti = 0, tq = 0;
        for (int b = 0; b < G.mod; ++b) {
            const int p = sf * G.mod + b;                 // interleaved position in the frame
            const int jj = p / G.I, ii = p - jj * G.I;
            const int src = (LDPC_N / G.I) * ii + jj;     // position in [info|parity]
            int bit;
            if (G.codeword) bit = G.codeword[src];
            else {
                const int8_t* ob = G.output_bits + (size_t)group * 32 * LDPC_N;
                bit = src < LDPC_K ? ob[frame * LDPC_K + src] : ob[32 * LDPC_K + frame * LDPC_M + (src - LDPC_K)];
            }
            const unsigned sh = half - (b >> 1) - 1;
            if (b & 1) tq += (unsigned)bit << sh;
            else ti += (unsigned)bit << sh;
        }
        re[k] = tab[ti];
        im[k] = tab[tq];
    }
}

// Same with InterleaveModType == 1 (identity interleaver, the shipped Profile.txt): the 2*MOD bits of the symbol pair are
// consecutive bytes starting at code bit 2*sp*MOD, fetched as 32-bit words (K, M and 2*MOD*sp are multiples of 4 and a
// pair never straddles the info / parity boundary because K is a multiple of 4, 8, 12 and 16).
template <int MOD>
__device__ __forceinline__ void map_symbol_pair_i1(const GenCore& G, int group, int frame, int sp, float (&re)[2], float (&im)[2]) {
    const int src0 = 2 * sp * MOD;
    const int8_t* p;
    if (G.codeword) p = G.codeword + src0;
    else {
        const int8_t* ob = G.output_bits + (size_t)group * 32 * LDPC_N;
        p = src0 < LDPC_K ? ob + frame * LDPC_K + src0 : ob + 32 * LDPC_K + frame * LDPC_M + (src0 - LDPC_K);
    }
    uint32_t w[MOD / 2];
#pragma unroll
    for (int i = 0; i < MOD / 2; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(p) + i);
    const float* tab = gray_table(MOD);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        unsigned ti = 0, tq = 0;
#pragma unroll
        for (int b = 0; b < MOD; ++b) {
            const int byte = k * MOD + b;
            const unsigned bit = (w[byte >> 2] >> (8 * (byte & 3))) & 1u;
            const unsigned sh = MOD / 2 - (b >> 1) - 1;
            if (b & 1) tq += bit << sh;
            else ti += bit << sh;
        }
        re[k] = tab[ti];
        im[k] = tab[tq];
    }
}

// ---- channel ---------------------------------------------------------------------------------------------------------
// Complex AWGN on the two symbols of pair sp.  One Philox call serves both symbols (two Box-Muller pairs): offset =
// symbol-pair index, subsequence = global frame index, so the stream does not depend on which kernel, GPU, stream or
// chunk produces the frame.  Box-Muller runs on the special-function unit (lg2 / sqrt / sin / cos approximations,
// |error| ~ 1e-6 relative): the producer is a Monte-Carlo noise source, its parity with the reference's 3-LCG stream is
// statistical by design, and the demapper / quantiser downstream are bit-exact on whatever symbols it emits.
__device__ __forceinline__ void add_awgn_pair(const GenCore& G, uint64_t gframe, int sp, float (&re)[2], float (&im)[2]) {
    uint32_t r[4];
    Philox::gen(G.seed, gframe, kNoiseStream + (uint64_t)sp, r);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float u1 = ((float)r[2 * k] + 0.5f) * 2.3283064365386963e-10f;  // (r+0.5)/2^32 in (0,1]
        const float th = ((float)r[2 * k + 1] * 2.3283064365386963e-10f - 0.5f) * 6.283185307179586f;  // [-pi, pi)
        float sq;
        asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(-1.3862943611198906f * __log2f(u1)));  // sqrt(-2 ln u1)
        const float rad = G.sigma_d * sq;
        float sn, cs;
        __sincosf(th, &sn, &cs);
        re[k] = __fadd_rn(__fmul_rn(rad, cs), re[k]);
        im[k] = __fadd_rn(__fmul_rn(rad, sn), im[k]);
    }
}

// Noisy symbols 2*sp and 2*sp+1 of one frame (general path).
__device__ __forceinline__ void gen_symbol_pair(const GenCore& G, int group, int frame, uint64_t gframe, int sp,
                                                float (&re)[2], float (&im)[2]) {
    map_symbol_pair(G, group, frame, sp, re, im);
    if (G.add_noise) add_awgn_pair(G, gframe, sp, re, im);
}
template <int MOD>
__device__ __forceinline__ void gen_symbol_pair_i1(const GenCore& G, int group, int frame, uint64_t gframe, int sp,
                                                   float (&re)[2], float (&im)[2]) {
    map_symbol_pair_i1<MOD>(G, group, frame, sp, re, im);
    if (G.add_noise) add_awgn_pair(G, gframe, sp, re, im);
}

// the 2*MOD quantised LLRs of a symbol pair, in transmitted-bit order
template <int MOD>
__device__ __forceinline__ void demap_quant_pair(const float (&re)[2], const float (&im)[2], float scale, int qbits, int (&q)[2 * MOD]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float llr[kMaxMod];
        demap_symbol(re[k], im[k], MOD, llr);
#pragma unroll
        for (int b = 0; b < MOD; ++b) q[k * MOD + b] = quant_cfg(llr[b], scale, qbits);
    }
}

}  // namespace ldpc
