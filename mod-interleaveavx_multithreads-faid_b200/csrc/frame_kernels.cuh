// frame_kernels.cuh -- frame generation and scoring kernels that feed / follow the decoders.
//
//   quantize_kernel        CLDPC::float2LimitChar_4bit                               CLDPC.cpp:4524-4582
//   demap_kernel           CModulate::Demodulation + AfterDeModulationDeInterleaver  CModulate.cpp:156-212,270-362
//                          (+ the quantiser, fused)
//   generate_kernel        BeforeModulationInterleaver + Modulation                  CModulate.cpp:95-152,216-264
//                          + CChannel::AWGNChannel (Philox4x32-10 instead of the     CChannel.cpp:71-97
//                            3-LCG/Box-Muller stream: statistical parity by design)
//                          + demap + de-interleave + quantise, one pass, only int8 LLRs reach HBM
//   encode_group_kernel    CLDPC::Encode (GenMatrix is empty in the reference; the   CLDPC.cpp:68-126
//                          systematic encoder is p = Hp^-1 (Hs s), Hp^-1 block-circulant)
//   info_bits_kernel       CLDPC::GenMsgSeq (rand()%2 -> Philox)                     CLDPC.cpp:60-66
//   count_errors_kernel    CLDPC::CalculateErrors                                    CLDPC.cpp:4819-4995
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode_kernels.cuh"
#include "ldpc_b200.h"

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
__constant__ float c_tab_qpsk[2] = {-0.707107f, 0.707107f};
__constant__ float c_tab_16qam[4] = {-0.316228f, -0.948683f, 0.316228f, 0.948683f};
__constant__ float c_tab_64qam[8] = {-0.462910f, -0.154303f, -0.771517f, -1.08012f, 0.462910f, 0.154303f, 0.771517f, 1.08012f};

// float2LimitChar_4bit on one value.  _mm256_cvttps_epi32 yields INT_MIN for NaN / |x| >= 2^31, which the
// saturating packs and the clamp turn into -7 (CLDPC.cpp:4555-4573).
__device__ __forceinline__ int quant4(float x, float scale) {
    const float p = __fmul_rn(x, scale);
    if (!(p >= -2147483648.0f && p < 2147483648.0f)) return -7;
    const int t = __float2int_rz(p);
    return max(-7, min(7, t));
}

__global__ void quantize_kernel(const float* __restrict__ in, int8_t* __restrict__ out, int64_t n, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = i; k < n; k += stride) out[k] = (int8_t)quant4(in[k], scale);
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
__device__ __forceinline__ void demap_symbol(float re, float im, int mod, float (&llr)[6]) {
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
    }
}

struct GenParams {
    const int8_t* output_bits;  // [groups][32*N] two-region layout, or nullptr with `codeword`
    const int8_t* codeword;     // [N] same codeword for every frame (FakeEncoder), or nullptr
    const float* symbols_in;    // demap-only mode: noisy symbols instead of map + noise
    float* symbols_out;         // optional
    float* llr_float;           // optional: DeInterLeaveSeq
    int8_t* fix;                // fixInput
    int n_groups, mod, I;
    float sigma_d;              // per real dimension: sigma / sqrt(2)
    float scale;
    uint64_t seed, first_frame;
    int add_noise;
};

// one thread per symbol
__global__ void generate_kernel(const GenParams P) {
    const int sym_per_frame = kN / P.mod;
    const int64_t total = (int64_t)P.n_groups * 32 * sym_per_frame;
    const int half = P.mod / 2;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gframe = s / sym_per_frame;  // frame index within the call
        const int sf = (int)(s - gframe * sym_per_frame);
        const int group = (int)(gframe >> 5), frame = (int)(gframe & 31);
        float re, im;
        if (P.symbols_in) {
            re = P.symbols_in[2 * s];
            im = P.symbols_in[2 * s + 1];
        } else {
            // interleave + Gray map (CModulate.cpp:137-149, 243-262)
            unsigned ti = 0, tq = 0;
            for (int b = 0; b < P.mod; ++b) {
                const int p = sf * P.mod + b;                 // interleaved position in the frame
                const int jj = p / P.I, ii = p - jj * P.I;
                const int src = (kN / P.I) * ii + jj;         // position in [info|parity]
                int bit;
                if (P.codeword) bit = P.codeword[src];
                else {
                    const int8_t* ob = P.output_bits + (size_t)group * 32 * kN;
                    bit = src < kK ? ob[frame * kK + src] : ob[32 * kK + frame * kM + (src - kK)];
                }
                const unsigned sh = half - (b >> 1) - 1;
                if (b & 1) tq += (unsigned)bit << sh;
                else ti += (unsigned)bit << sh;
            }
            const float* tab = P.mod == 2 ? c_tab_qpsk : P.mod == 4 ? c_tab_16qam : c_tab_64qam;
            re = tab[ti];
            im = tab[tq];
            if (P.add_noise) {
                uint32_t r[4];
                Philox::gen(P.seed, P.first_frame + (uint64_t)gframe, kNoiseStream + (uint64_t)sf, r);
                // Box-Muller on (0,1] x [0,1): both outputs are used (the reference throws the sine away)
                const float u1 = ((float)r[0] + 0.5f) * 2.3283064365386963e-10f;  // (r+0.5)/2^32, never 0
                const float u2 = (float)r[1] * 2.3283064365386963e-10f;
                const float rad = P.sigma_d * sqrtf(-2.0f * logf(u1));
                float sn, cs;
                sincospif(2.0f * u2, &sn, &cs);
                re = __fadd_rn(__fmul_rn(rad, cs), re);
                im = __fadd_rn(__fmul_rn(rad, sn), im);
            }
        }
        if (P.symbols_out) {
            P.symbols_out[2 * s] = re;
            P.symbols_out[2 * s + 1] = im;
        }
        float llr[6];
        demap_symbol(re, im, P.mod, llr);
        int8_t* fix = P.fix ? P.fix + (size_t)group * 32 * kN : nullptr;
        float* lf = P.llr_float ? P.llr_float + (size_t)group * 32 * kN : nullptr;
        for (int b = 0; b < P.mod; ++b) {
            const int off = deint_offset(frame, sf * P.mod + b, P.I);
            if (lf) lf[off] = llr[b];
            if (fix) fix[off] = (int8_t)quant4(llr[b], P.scale);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// info bits + systematic encoder, one CTA per group, frame-sliced (bit f of a word = frame f)
// ---------------------------------------------------------------------------------------------------------
__global__ void info_bits_kernel(int8_t* __restrict__ input_bits, int n_groups, uint64_t seed, uint64_t first_frame) {
    // 128 bits per Philox call; thread handles 128 consecutive info bits of one frame (K = 114 * 128)
    const int per_frame = kK / 128;
    const int64_t total = (int64_t)n_groups * 32 * per_frame;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gframe = w / per_frame;
        const int q = (int)(w - gframe * per_frame);
        uint32_t r[4];
        Philox::gen(seed, first_frame + (uint64_t)gframe, kInfoStream + (uint64_t)q, r);
        int8_t* o = input_bits + (size_t)gframe * kK + (size_t)q * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            for (int b = 0; b < 32; b += 4) {
                const uint32_t nib = (r[k] >> b) & 0xFu;
                *reinterpret_cast<uint32_t*>(o + k * 32 + b) = (nib * 0x00204081u) & 0x01010101u;
            }
    }
}

constexpr int kEncThreads = 1024;

// input_bits int8 [32][K] -> output_bits int8 two-region [32*K | 32*M]
__global__ void __launch_bounds__(kEncThreads, 1) encode_group_kernel(const int8_t* __restrict__ input_bits,
                                                                      int8_t* __restrict__ output_bits, int n_groups) {
    extern __shared__ uint32_t esm[];
    uint32_t* sbits = esm;       // [K] frame-sliced info bits
    uint32_t* tvec = esm + kK;   // [M] frame-sliced Hs*s
    __shared__ uint32_t tile[32][33];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x;
    const int8_t* in = input_bits + (size_t)g * 32 * kK;
    int8_t* out = output_bits + (size_t)g * 32 * kN;

    // 32x32 bit transpose per 32 consecutive info positions; also copies the systematic part through
    for (int blk = warp; blk < kK / 32; blk += kEncThreads / 32) {
        const int j0 = blk * 32;
        for (int f = 0; f < 32; ++f) {
            const int8_t v = in[f * kK + j0 + lane];
            out[f * kK + j0 + lane] = v;
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, v & 1);
            if (lane == 0) tile[warp][f] = b;
        }
        __syncwarp();
        uint32_t w = 0;
        for (int f = 0; f < 32; ++f) w |= ((tile[warp][f] >> lane) & 1u) << f;
        sbits[j0 + lane] = w;
        __syncwarp();
    }
    __syncthreads();
    // t = Hs * s : row (l, r) XORs the info-part entries of its layer
    for (int row = tid; row < kM; row += kEncThreads) {
        const int l = row >> 8, r = row & 255;
        uint32_t x = 0;
        for (int e = c_code.layer_start[l]; e < c_code.layer_start[l + 1]; ++e) {
            const int c = c_code.circ_col[e];
            if (c < kK / kZ) x ^= sbits[c * 256 + ((c_code.circ_shift[e] + r) & 255)];
        }
        tvec[row] = x;
    }
    __syncthreads();
    // p = Hp^-1 * t with Hp^-1[i*256+r][j*256+c] = q_ij[(r - c) mod 256]
    for (int row = tid; row < kM; row += kEncThreads) {
        const int i = row >> 8, r = row & 255;
        uint32_t x = 0;
        for (int j = 0; j < LDPC_MB; ++j) {
            const uint32_t* t = tvec + j * 256;
#pragma unroll 1
            for (int w = 0; w < 8; ++w) {
                uint32_t q = c_code.hpinv[i][j][w];
                while (q) {
                    const int k = 32 * w + __ffs(q) - 1;
                    q &= q - 1;
                    x ^= t[(r - k) & 255];
                }
            }
        }
        // scatter the 32 frames' parity bit of this row
        for (int f = 0; f < 32; ++f) out[32 * kK + f * kM + row] = (int8_t)((x >> f) & 1u);
    }
}

// ---------------------------------------------------------------------------------------------------------
// CalculateErrors: info bits only (CLDPC.cpp:4847-4876).  One warp per frame.
// counters: [0] TestFrame [1] ErrorFrame [2] ErrorBits [3] LT3ErrBitFrame
// ---------------------------------------------------------------------------------------------------------
__global__ void count_errors_kernel(const int8_t* __restrict__ input_bits, const int8_t* __restrict__ decoded,
                                    int n_frames, unsigned long long* __restrict__ counters, int in_stride) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int f = warp; f < n_frames; f += nwarps) {
        const uint4* a = reinterpret_cast<const uint4*>(input_bits + (size_t)f * in_stride);  // stride 0: one fixed codeword
        const uint4* d = reinterpret_cast<const uint4*>(decoded + (size_t)f * kN);
        int eb = 0;
        for (int q = lane; q < kK / 16; q += 32) {
            const uint4 x = a[q], y = d[q];
            // bytes are 0/1; differing bytes have their lowest bit set after the xor
            eb += __popc((x.x ^ y.x) & 0x01010101u) + __popc((x.y ^ y.y) & 0x01010101u) +
                  __popc((x.z ^ y.z) & 0x01010101u) + __popc((x.w ^ y.w) & 0x01010101u);
        }
        eb = __reduce_add_sync(0xFFFFFFFFu, eb);
        if (lane == 0) {
            atomicAdd(&counters[0], 1ull);
            if (eb > 0) {
                atomicAdd(&counters[1], 1ull);
                atomicAdd(&counters[2], (unsigned long long)eb);
                if (eb < 3) atomicAdd(&counters[3], 1ull);
            }
        }
    }
}

// histograms of the per-group iteration counts (iterCount.txt of the reference + min-sum iterations)
__global__ void group_hist_kernel(const int32_t* __restrict__ bf_iters, const int32_t* __restrict__ its, int n_groups,
                                  unsigned long long* __restrict__ counters) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    atomicAdd(&counters[LDPC_B200_CNT_GROUPS], 1ull);
    atomicAdd(&counters[LDPC_B200_CNT_MS_ITERS_SUM], (unsigned long long)its[g]);
    atomicAdd(&counters[LDPC_B200_CNT_BF_HIST + min(max(bf_iters[g], 0), 50)], 1ull);
    atomicAdd(&counters[LDPC_B200_CNT_MS_HIST + min(max(its[g], 0), 63)], 1ull);
}

// ---------------------------------------------------------------------------------------------------------
struct FrameState {
    cudaStream_t stream = nullptr;
    unsigned long long* d_counters = nullptr;  // [LDPC_B200_NUM_COUNTERS]
    // simulate() workspace, sized for sim_groups groups
    int sim_groups = 0;
    int8_t *d_info = nullptr, *d_tx = nullptr, *d_fix = nullptr, *d_dec = nullptr, *d_codeword = nullptr;
    // staging for host-pointer calls of the small entry points
    void* d_tmp[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t tmp_bytes[4] = {0, 0, 0, 0};
    // NCCL (resolved lazily with dlopen)
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
};

inline int frame_state_init(FrameState& fs, const ldpc_b200_config&) {
    if (cudaStreamCreateWithFlags(&fs.stream, cudaStreamNonBlocking) != cudaSuccess) return LDPC_B200_ECUDA;
    if (cudaMalloc(&fs.d_counters, LDPC_B200_NUM_COUNTERS * sizeof(unsigned long long)) != cudaSuccess) return LDPC_B200_ENOMEM;
    if (cudaMalloc(&fs.d_codeword, kN) != cudaSuccess) return LDPC_B200_ENOMEM;
    return 0;
}
inline void frame_state_free(FrameState& fs) {
    for (auto& p : fs.d_tmp)
        if (p) cudaFree(p);
    if (fs.d_info) cudaFree(fs.d_info);
    if (fs.d_tx) cudaFree(fs.d_tx);
    if (fs.d_fix) cudaFree(fs.d_fix);
    if (fs.d_dec) cudaFree(fs.d_dec);
    if (fs.d_codeword) cudaFree(fs.d_codeword);
    if (fs.d_counters) cudaFree(fs.d_counters);
    if (fs.stream) cudaStreamDestroy(fs.stream);
    fs = FrameState();
}

}  // namespace ldpc
