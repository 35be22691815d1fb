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

#include "decode_kernels.cuh"  // brings gen_device.cuh
#include "ldpc_b200.h"

namespace ldpc {

__global__ void quantize_kernel(const float* __restrict__ in, int8_t* __restrict__ out, int64_t n, float scale, int bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = i; k < n; k += stride) out[k] = (int8_t)quant_cfg(in[k], scale, bits);
}


struct GenParams {
    GenCore core;               // what defines the transmitted symbols and the noise (shared with the fused decoder loader)
    const float* symbols_in;    // demap-only mode: noisy symbols instead of map + noise
    float* symbols_out;         // optional
    float* llr_float;           // optional: DeInterLeaveSeq
    int8_t* fix;                // fixInput
    int n_groups;
};

// InterleaveModType == 1, int8 LLRs only: word-wise bit fetch, 4 / 8 / 12 LLR bytes stored as 32-bit words
template <int MOD>
__global__ void generate_i1_kernel(const GenParams P) {
    const GenCore& G = P.core;
    constexpr int pairs_per_frame = kN / MOD / 2;
    const int64_t total = (int64_t)P.n_groups * 32 * pairs_per_frame;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gframe = s / pairs_per_frame;
        const int sp = (int)(s - gframe * pairs_per_frame);
        const int group = (int)(gframe >> 5), frame = (int)(gframe & 31);
        float re[2], im[2];
        int q[2 * MOD];
        gen_symbol_pair_i1<MOD>(G, tx_group_of(G, group), frame, G.first_frame + (uint64_t)gframe, sp, re, im);
        demap_quant_pair<MOD>(re, im, G.scale, G.qbits, q);
        const int dst0 = 2 * MOD * sp;
        int8_t* o = P.fix + (size_t)group * 32 * kN + (dst0 < kK ? frame * kK + dst0 : 32 * kK + frame * kM + (dst0 - kK));
#pragma unroll
        for (int i = 0; i < MOD / 2; ++i)
            reinterpret_cast<uint32_t*>(o)[i] = (uint32_t)(q[4 * i] & 0xFF) | ((uint32_t)(q[4 * i + 1] & 0xFF) << 8) |
                                                ((uint32_t)(q[4 * i + 2] & 0xFF) << 16) | ((uint32_t)(q[4 * i + 3] & 0xFF) << 24);
    }
}

// general path: one thread per PAIR of symbols (one Philox call yields the four normals of two symbols)
__global__ void generate_kernel(const GenParams P) {
    const GenCore& G = P.core;
    const int pairs_per_frame = kN / G.mod / 2;
    const int64_t total = (int64_t)P.n_groups * 32 * pairs_per_frame;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gframe = s / pairs_per_frame;  // frame index within the call
        const int sp = (int)(s - gframe * pairs_per_frame);
        const int group = (int)(gframe >> 5), frame = (int)(gframe & 31);
        float re[2], im[2];
        if (P.symbols_in) {
            const float4 v = reinterpret_cast<const float4*>(P.symbols_in)[s];
            re[0] = v.x; im[0] = v.y; re[1] = v.z; im[1] = v.w;
        } else {
            gen_symbol_pair(G, tx_group_of(G, group), frame, G.first_frame + (uint64_t)gframe, sp, re, im);
        }
        if (P.symbols_out) reinterpret_cast<float4*>(P.symbols_out)[s] = make_float4(re[0], im[0], re[1], im[1]);
        int8_t* fix = P.fix ? P.fix + (size_t)group * 32 * kN : nullptr;
        float* lf = P.llr_float ? P.llr_float + (size_t)group * 32 * kN : nullptr;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float llr[kMaxMod];
            demap_symbol(re[k], im[k], G.mod, llr);
            for (int b = 0; b < G.mod; ++b) {
                const int off = deint_offset(frame, (2 * sp + k) * G.mod + b, G.I);
                if (lf) lf[off] = llr[b];
                if (fix) fix[off] = (int8_t)quant_cfg(llr[b], G.scale, G.qbits);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// info bits + systematic encoder, one CTA per group, frame-sliced (bit f of a word = frame f)
// ---------------------------------------------------------------------------------------------------------
// group_stride: group k of the output is the info block of global group (first_frame / 32 + k * group_stride), i.e. with
// codeword reuse only every group_stride-th group's bits are ever drawn.
__global__ void info_bits_kernel(int8_t* __restrict__ input_bits, int n_groups, uint64_t seed, uint64_t first_frame,
                                 int group_stride) {
    // 128 bits per Philox call; thread handles 128 consecutive info bits of one frame (K = 114 * 128)
    const int per_frame = kK / 128;
    const int64_t total = (int64_t)n_groups * 32 * per_frame;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gframe = w / per_frame;
        const int q = (int)(w - gframe * per_frame);
        uint32_t r[4];
        const uint64_t sub = first_frame + (uint64_t)(gframe >> 5) * 32ull * (uint64_t)group_stride + (uint64_t)(gframe & 31);
        Philox::gen(seed, sub, kInfoStream + (uint64_t)q, r);
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
            if (blockIdx.y == 0) out[f * kK + j0 + lane] = v;
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
    if (gridDim.y == 1) {
        // throughput shape (many groups): one CTA does all 3072 parity rows of its group
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
    } else {
        // latency shape (few groups, e.g. one encode per 50 noise blocks): 12 CTAs per group, CTA y owns parity block row y,
        // four threads share a row (three column blocks each) and meet in shared memory
        const int i = blockIdx.y, r = tid & 255, part = tid >> 8;
        uint32_t x = 0;
        for (int j = part; j < LDPC_MB; j += 4) {
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
        __syncthreads();               // every thread is done reading sbits: reuse it for the partial sums
        sbits[part * 256 + r] = x;
        __syncthreads();
        if (part == 0) {
            x = sbits[r] ^ sbits[256 + r] ^ sbits[512 + r] ^ sbits[768 + r];
            const int row = i * 256 + r;
            for (int f = 0; f < 32; ++f) out[32 * kK + f * kM + row] = (int8_t)((x >> f) & 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// CalculateErrors: info bits only (CLDPC.cpp:4847-4876).  One warp per frame.
// counters: [0] TestFrame [1] ErrorFrame [2] ErrorBits [3] LT3ErrBitFrame
// ---------------------------------------------------------------------------------------------------------
__global__ void count_errors_kernel(const int8_t* __restrict__ input_bits, const int8_t* __restrict__ decoded,
                                    int n_frames, unsigned long long* __restrict__ counters, int in_stride,
                                    int tx_reuse, uint32_t tx_c0, uint64_t group0) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int f = warp; f < n_frames; f += nwarps) {
        // stride 0: one fixed codeword; tx_reuse > 1: the frame's info bits live in the (shared) codeword group
        const size_t fi = tx_reuse > 1 ? (size_t)((group0 + (uint64_t)(f >> 5)) / (uint64_t)tx_reuse - tx_c0) * 32 + (f & 31) : (size_t)f;
        const uint4* a = reinterpret_cast<const uint4*>(input_bits + fi * in_stride);
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
