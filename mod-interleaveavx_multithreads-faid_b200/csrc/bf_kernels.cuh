// bf_kernels.cuh -- group finalisation: early-stop resolution, bit-flipping post-processing, output formatting.
//
// Replaces, per group of 32 frames:
//   * the group-level early stop of the min-sum loops       CDecoder_OMS.cpp:325-327, CDecoder_FAID.cpp:616-618
//   * plain BF            (DecodeMethod 3)                  CDecoder_OMSBF.cpp:2959-3511
//   * DTBF                (DecodeMethod 2 and 4)            CDecoder_FAID.cpp:6411-7088, CDecoder_OMS_DTBF.cpp:2968-3650
//   * 2-bit DTBF "2B1C"   (DecodeMethod 5)                  CDecoder_FAID_2B1C.cpp:6124-6813
//   * the hard decision + inverse transpose into decodedBits  CTool.cpp:291-575, CLDPC.cpp:2268-2270
//
// One CTA per group, one warp per frame.  All BF state is bit-packed along the circulant dimension: a word holds
// 32 consecutive positions of one 256-bit block column, so a circulant permutation is a 256-bit rotation done
// with one funnel shift per word, the syndrome is an XOR of 22/23 rotated words, and the per-bit vote counts are
// carry-save bit planes.  The 32 frames only meet in __syncthreads_or (the reference's group-level `break`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "decode_kernels.cuh"

namespace ldpc {

struct FinParams {
    const uint32_t* final_hard;
    const uint32_t* snap;
    const uint32_t* grp_cnt;
    const int32_t* first_zero;
    int n_groups, max_iter, planes, has_syndrome;
    int bf_mode, bf_max_iter, L0, L1, delta, alpha, rcw;
    int fast_bf;            // unrolled bit-flipping stage (rcw == 3, alpha <= 1); 0 = generic table-driven loops
    unsigned long long* dbg; // LDPC_DEBUG_BOUNDS builds: violation record (see decode_kernels.cuh)
    int wpf;                // shared-memory words per frame (fin_layout_words)
    int unsat_bufs;         // 1, or 2: the unrolled BF / DTBF stage keeps the syndrome up to date incrementally (ping-pong)
    int8_t* decoded;        // reference layout: int8 [group][32][N], or nullptr
    uint32_t* hard_packed;  // native layout: [frame][kHW], or nullptr
    int32_t* bf_iters;      // [groups] or nullptr
    int32_t* its_per_group; // [groups] or nullptr
    int32_t* conv_iter;     // [frames] or nullptr
};

enum { BF_NONE = 0, BF_PLAIN = 1, BF_DTBF = 2, BF_2B1C = 3 };

constexpr int kFinThreads = 1024;
constexpr int kUnsatW = LDPC_MB * 8;  // 96 words of row-unsatisfied flags per frame

// Shared-memory words per frame of finalize_kernel: hard | unsat x unsat_bufs | diff (DTBF, 2B1C) | hard2 (2B1C), padded so
// that wpf mod 32 is 8 or 24: a warp of the unrolled stage touches 4 frames x 8 consecutive words per access, which then
// fall into 32 different banks (the unpadded DTBF layout, 1200 words, put frames 0 / 2 and 1 / 3 on the same banks).
// The second syndrome buffer does not fit beside the four arrays of 2B1C (227 KB), which therefore recomputes.
__host__ __device__ constexpr int fin_unsat_bufs(int bf_mode, bool do_bf) { return (do_bf && (bf_mode == 1 || bf_mode == 2)) ? 2 : 1; }
__host__ __device__ constexpr int fin_layout_words(int bf_mode, bool do_bf) {
    int w = kHW;
    if (do_bf) w += kUnsatW * fin_unsat_bufs(bf_mode, do_bf) + (bf_mode != 1 ? kHW : 0) + (bf_mode == 3 ? kHW : 0);
    while (w % 32 != 8 && w % 32 != 24) w += 8;
    return w;
}

__device__ __forceinline__ int sat8i(int x) { return x > 127 ? 127 : (x < -128 ? -128 : x); }

// v (5 bit planes) += bit
__device__ __forceinline__ void planes_add_bit(uint32_t (&v)[5], uint32_t bit) {
    uint32_t c = bit;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const uint32_t t = v[b] & c;
        v[b] ^= c;
        c = t;
    }
}
// v += k on the positions selected by mask (k small, non-negative)
__device__ __forceinline__ void planes_add_const(uint32_t (&v)[5], uint32_t mask, int k) {
#pragma unroll
    for (int b0 = 0; b0 < 5; ++b0) {
        if ((k >> b0) & 1) {
            uint32_t c = mask;
#pragma unroll
            for (int b = b0; b < 5; ++b) {
                const uint32_t t = v[b] & c;
                v[b] ^= c;
                c = t;
            }
        }
    }
}
// positions where v >= T (T scalar)
__device__ __forceinline__ uint32_t planes_ge(const uint32_t (&v)[5], int T) {
    if (T <= 0) return 0xFFFFFFFFu;
    if (T > 31) return 0u;
    uint32_t gt = 0, eq = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 4; b >= 0; --b) {
        const uint32_t tb = ((T >> b) & 1) ? 0xFFFFFFFFu : 0u;
        gt |= eq & v[b] & ~tb;
        eq &= ~(v[b] ^ tb);
    }
    return gt | eq;
}

// vote count planes of the 32 code bits [32*i, 32*i+32) of block column c: number of unsatisfied rows touching each
__device__ __forceinline__ void vote_planes(uint32_t (&v)[5], const uint32_t* unsat, int c, int i,
                                            const uint16_t* s_col_start, const uint8_t* s_col_layer,
                                            const uint8_t* s_col_lshift) {
#pragma unroll
    for (int b = 0; b < 5; ++b) v[b] = 0;
    for (int e = s_col_start[c]; e < s_col_start[c + 1]; ++e) {
        const int l = s_col_layer[e], s = s_col_lshift[e];
        const int p = (32 * i - s) & 255;
        const uint32_t* u = unsat + l * 8;
        planes_add_bit(v, __funnelshift_r(u[p >> 5], u[((p >> 5) + 1) & 7], p & 31));
    }
}


// ---- unrolled bit-flipping stage ---------------------------------------------------------------------------------------
// Thread mapping: warp = (quarter << 3) | quad; the 4 warps of a quad own frames 4*quad .. 4*quad+3 and split the layer /
// column tasks by (index & 3) == quarter; lane = (frame_sub << 3) | word, word = which 32-bit word of a 256-bit block
// column.  With that mapping block columns, layers and shifts are literals (X-macros of ldpc_code_tables.h): a rotated
// word is two LDS at fixed offsets from eight precomputed per-lane pointers and one funnel shift.
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
    s = a ^ b ^ c;
    cy = (a & b) | (a & c) | (b & c);
}
// 32 consecutive bits starting at bit (32 * word + S) mod 256 of the 256-bit string at base[0..8); ptr[a] = base + ((word + a) & 7)
template <int S>
__device__ __forceinline__ uint32_t rot_word(const uint32_t* const (&ptr)[8], int off) {
    constexpr int a = (S >> 5) & 7, b = S & 31;
    if (b == 0) return ptr[a][off];
    return __funnelshift_r(ptr[a][off], ptr[(a + 1) & 7][off], b);
}
// vote planes (v0 = 1s, v1 = 2s, v2 = 4s, v3 = 8s) of W one-bit inputs
template <int W>
__device__ __forceinline__ void count_planes(const uint32_t (&x)[12], uint32_t (&v)[4]) {
    static_assert(W == 3 || W == 6 || W == 11 || W == 12, "column weights of the 50G-PON code");
    if (W == 3) {
        full_add(x[0], x[1], x[2], v[0], v[1]);
        v[2] = v[3] = 0;
    } else if (W == 6) {
        uint32_t s1, c1, s2, c2;
        full_add(x[0], x[1], x[2], s1, c1);
        full_add(x[3], x[4], x[5], s2, c2);
        v[0] = s1 ^ s2;
        full_add(c1, c2, s1 & s2, v[1], v[2]);
        v[3] = 0;
    } else {  // 11 or 12 (x[11] = 0 for 11)
        uint32_t s1, c1, s2, c2, s3, c3, s4, c4, s5, c5, t1, d1, t2, d2;
        full_add(x[0], x[1], x[2], s1, c1);
        full_add(x[3], x[4], x[5], s2, c2);
        full_add(x[6], x[7], x[8], s3, c3);
        full_add(x[9], x[10], x[11], s4, c4);
        full_add(s1, s2, s3, s5, c5);
        v[0] = s5 ^ s4;
        full_add(c1, c2, c3, t1, d1);
        full_add(c4, c5, s5 & s4, t2, d2);
        v[1] = t1 ^ t2;
        full_add(d1, d2, t1 & t2, v[2], v[3]);
    }
}
// positions with count >= T, T in 1..5 given as a lane-varying scalar
__device__ __forceinline__ uint32_t planes4_ge(const uint32_t (&v)[4], int T) {
    const uint32_t hi = v[2] | v[3];
    const uint32_t ge1 = v[0] | v[1] | hi, ge2 = v[1] | hi, ge3 = (v[1] & v[0]) | hi, ge4 = hi, ge5 = v[3] | (v[2] & (v[1] | v[0]));
    return T <= 1 ? ge1 : T == 2 ? ge2 : T == 3 ? ge3 : T == 4 ? ge4 : T == 5 ? ge5 : 0u;
}

#define LDPC_BF_SYN_EDGE(j, c, s, w) X ^= rot_word<(s)>(hp, (c) * 8);
#define LDPC_BF_SYN_LAYER(LY)                                   \
    if (quarter == ((LY) & 3)) {                                \
        uint32_t X = 0;                                         \
        LDPC_EDGES_L##LY(LDPC_BF_SYN_EDGE)                      \
        ucur[(LY) * 8 + word] = X;                              \
    }
// vote of row (layer, r) reaches code bit (shift + r) mod 256: the unsat string rotated by 256 - shift
#define LDPC_BF_VOTE_EDGE(k, l, s) x[k] = rot_word<((256 - (s)) & 255)>(up, (l) * 8);
// own syndrome words of this lane (layers LY with LY % 4 == quarter): OR into `any`, copy into the other buffer
#define LDPC_BF_ANY_LAYER(LY)                                   \
    if (quarter == ((LY) & 3)) {                                \
        const uint32_t w = ucur[(LY) * 8 + word];               \
        any |= w;                                               \
        if (incr) uoth[(LY) * 8 + word] = w;                    \
    }
// Incremental syndrome: flipping the code bits `fl` of word `word` of a block column toggles, in every layer (l, shift s) of
// that column, the rows (v - s) mod 256 -- the same rotation the vote of that edge reads, so the mask lands at bit
// 32 (word + a) + b of the layer's 256-bit string with (a, b) = ((256 - s) >> 5, (256 - s) & 31).  H (hard ^ fl) = H hard ^ H fl.
#define LDPC_BF_UPD_EDGE(k, l, s)                                                                     \
    {                                                                                                 \
        constexpr int a_ = (((256 - (s)) & 255) >> 5), b_ = ((256 - (s)) & 31);                       \
        atomicXor(&uoth[(l) * 8 + ((word + a_) & 7)], fl_ << b_);                                     \
        if (b_) atomicXor(&uoth[(l) * 8 + ((word + a_ + 1) & 7)], fl_ >> ((32 - b_) & 31));           \
    }

__global__ void __launch_bounds__(kFinThreads, 1) finalize_kernel(const FinParams P) {
    extern __shared__ uint32_t sm[];
    __shared__ uint8_t s_ecol[LDPC_NCIRC], s_eshift[LDPC_NCIRC], s_col_layer[LDPC_NCIRC], s_col_lshift[LDPC_NCIRC];
    __shared__ uint16_t s_lstart[LDPC_MB + 1], s_col_start[LDPC_NB + 1];
    __shared__ uint8_t s_colw[LDPC_NB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x;
    const int frame = g * 32 + warp;

    for (int i = tid; i < LDPC_NCIRC; i += kFinThreads) {
        s_ecol[i] = c_code.circ_col[i];
        s_eshift[i] = c_code.circ_shift[i];
        s_col_layer[i] = c_code.col_layer[i];
        s_col_lshift[i] = c_code.col_lshift[i];
    }
    if (tid <= LDPC_MB) s_lstart[tid] = c_code.layer_start[tid];
    if (tid <= LDPC_NB) s_col_start[tid] = c_code.col_start[tid];
    if (tid < LDPC_NB) s_colw[tid] = c_code.col_weight[tid];

    // ---- resolve the group's stop iteration ----
    int jstar = -1;
    if (P.has_syndrome) {
        const uint32_t* cnt = P.grp_cnt + (size_t)g * P.max_iter;
        for (int j = 0; j < P.max_iter; ++j)
            if (cnt[j] == 32u) { jstar = j; break; }
    }
    const int its = jstar >= 0 ? jstar : P.max_iter;
    if (P.conv_iter && lane == 0) {
        int cv = -1;
        if (P.has_syndrome) {
            const int m = P.first_zero[frame];
            const int lim = jstar >= 0 ? jstar : P.max_iter - 1;
            if (m && m - 1 <= lim) cv = m - 1;
        }
        P.conv_iter[frame] = cv;
    }
    LDPC_CHECK(P.dbg, jstar < P.max_iter && warp < 32 && g < P.n_groups && (size_t)(warp + 1) * P.wpf * 4 <= (size_t)32 * P.wpf * 4, DBG_FIN, g);
    const uint32_t* src = jstar >= 0 ? P.snap + (((size_t)frame * P.max_iter + jstar) * P.planes) * kHW
                                     : P.final_hard + (size_t)frame * P.planes * kHW;

    const bool do_bf = P.bf_mode != BF_NONE && P.bf_max_iter > 0;
    // per-frame smem slices
    const int words_per_frame = P.wpf;
    uint32_t* hard = sm + (size_t)warp * words_per_frame;
    uint32_t* unsat = hard + kHW;
    uint32_t* diff = unsat + kUnsatW * P.unsat_bufs;   // hard ^ hard_ch (DTBF / 2B1C)
    uint32_t* hard2 = diff + kHW;       // 2B1C second bit

    for (int w = lane; w < kHW; w += 32) {
        hard[w] = src[w];
        if (do_bf && P.bf_mode != BF_PLAIN) diff[w] = 0;
        if (do_bf && P.bf_mode == BF_2B1C) hard2[w] = src[kHW + w];
    }
    __syncthreads();

    int BFiter = 0;
    if (do_bf) {
        int t_prev = 1, Th = P.rcw, l0 = 0, l1 = 0;
        const int L0 = (int8_t)P.L0, L1 = (int8_t)P.L1;
        if (P.fast_bf) {
            // ---------------- unrolled stage (see the mapping comment above) ----------------
            __shared__ int s_flag[2][32];   // per frame: "some bit flipped" / plain BF: mask of reached vote levels
            const int quad = warp & 7, quarter = warp >> 3, fsub = lane >> 3, word = lane & 7;
            // Frames whose hard decisions already satisfy H when the stage starts are INERT: zero syndrome and no bit flipped
            // so far means zero votes (+ alpha * 0), which no threshold >= 1 reaches, in this and every later iteration -- but
            // the group keeps iterating as long as any of its 32 frames fails (CDecoder_FAID.cpp:6782-6784).  In the waterfall
            // that is most frames of most groups, so the lanes are re-dealt: the failing frames first, four to a quad of
            // warps, and quads left with inert frames only just keep the barriers company.
            __shared__ unsigned int s_active;
            if (tid == 0) s_active = 0u;
            if (tid < 64) (&s_flag[0][0])[tid] = 0;
            __syncthreads();
            int f = quad * 4 + fsub;  // frame of this lane (natural order for the first syndrome)
            uint32_t* hardF = sm + (size_t)f * words_per_frame;
            uint32_t* unsatF = hardF + kHW;
            // syndrome kept up to date from the flips (two buffers: votes read `ucur` while the flips update `uoth`)
            const bool incr = P.unsat_bufs == 2;
            uint32_t* ucur = unsatF;
            const uint32_t* hp[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) hp[a] = hardF + ((word + a) & 7);
            {
                LDPC_FOR_EACH_LAYER(LDPC_BF_SYN_LAYER)  // full syndrome H * hard of every frame, once
                uint32_t any0 = 0;
#define LDPC_BF_ANY0_LAYER(LY) if (quarter == ((LY) & 3)) any0 |= ucur[(LY) * 8 + word];
                LDPC_FOR_EACH_LAYER(LDPC_BF_ANY0_LAYER)
#undef LDPC_BF_ANY0_LAYER
                if (any0) atomicOr(&s_active, 1u << f);
            }
            __syncthreads();
            // Lane group `fsub` of every quad takes its frame from residue class f % 4 == fsub (8 frames each), failing ones
            // first: a quad then still holds one frame of every class, and with wpf % 32 in {8, 24} the four frames of a warp
            // access stay on four different bank groups, exactly as in natural order.
            const unsigned int act = s_active;
            const unsigned int cls = 0x11111111u << fsub;
            const int n_cls = __popc(act & cls);
            f = quad < n_cls ? (int)__fns(act & cls, 0, quad + 1) : (int)__fns(~act & cls, 0, quad - n_cls + 1);
            const int n_max = max(max(__popc(act & 0x11111111u), __popc(act & 0x22222222u)), max(__popc(act & 0x44444444u), __popc(act & 0x88888888u)));
            const bool quad_active = quad < n_max;  // warp-uniform
            hardF = sm + (size_t)f * words_per_frame;
            unsatF = hardF + kHW;
            uint32_t* diffF = unsatF + kUnsatW * P.unsat_bufs;
            uint32_t* hard2F = diffF + kHW;
            ucur = unsatF;
            uint32_t* uoth = unsatF + (incr ? kUnsatW : 0);
            const uint32_t* up[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                hp[a] = hardF + ((word + a) & 7);
                up[a] = unsatF + ((word + a) & 7);
            }
            while (BFiter < P.bf_max_iter) {
                uint32_t any = 0;
                if (quad_active) {
                    if (!incr && BFiter > 0) { LDPC_FOR_EACH_LAYER(LDPC_BF_SYN_LAYER) }  // 2B1C: recomputed (no room for a second buffer)
                    LDPC_FOR_EACH_LAYER(LDPC_BF_ANY_LAYER)
                }
                if (!__syncthreads_or(any != 0)) break;  // group-level break (CDecoder_FAID.cpp:6782-6784)
                int* flag = &s_flag[BFiter & 1][f];
                if (P.bf_mode == BF_PLAIN) {
                    // CDecoder_OMSBF.cpp:2994,3327-3335: flip every bit with votes >= min(max(1, max votes), 5)
                    uint32_t lv = 0;  // bit k: some bit of this frame has >= k votes
#define LDPC_BF_PLAIN_MAX(C)                                                                             \
    if (quarter == ((C) & 3)) {                                                                          \
        uint32_t x[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, v[4];                                      \
        LDPC_COL_EDGES_C##C(LDPC_BF_VOTE_EDGE)                                                           \
        count_planes<LDPC_COLW_C##C>(x, v);                                                              \
        const uint32_t hi = v[2] | v[3];                                                                 \
        lv |= ((v[1] | hi) ? 4u : 0u) | (((v[1] & v[0]) | hi) ? 8u : 0u) | (hi ? 16u : 0u) |             \
              ((v[3] | (v[2] & (v[1] | v[0]))) ? 32u : 0u);                                              \
    }
                    if (quad_active) {
                        LDPC_FOR_EACH_COL(LDPC_BF_PLAIN_MAX)
                        if (lv) atomicOr(flag, (int)lv);
                    }
#undef LDPC_BF_PLAIN_MAX
                    __syncthreads();
                    const int m = *flag;
                    const int thr = (m & 32) ? 5 : (m & 16) ? 4 : (m & 8) ? 3 : (m & 4) ? 2 : 1;
#define LDPC_BF_PLAIN_FLIP(C)                                                                            \
    if (quarter == ((C) & 3)) {                                                                          \
        uint32_t x[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, v[4];                                      \
        LDPC_COL_EDGES_C##C(LDPC_BF_VOTE_EDGE)                                                           \
        count_planes<LDPC_COLW_C##C>(x, v);                                                              \
        const uint32_t fl_ = planes4_ge(v, thr);                                                         \
        if (fl_) {                                                                                       \
            hardF[(C) * 8 + word] ^= fl_;                                                                \
            if (incr) { LDPC_COL_EDGES_C##C(LDPC_BF_UPD_EDGE) }                                          \
        }                                                                                                \
    }
                    if (quad_active) { LDPC_FOR_EACH_COL(LDPC_BF_PLAIN_FLIP) }
#undef LDPC_BF_PLAIN_FLIP
                    __syncthreads();  // the flips (and the last reads of unsat) precede the next syndrome
                } else {
                    // threshold automaton, per frame (CDecoder_FAID.cpp:6787-6799)
                    if (!t_prev) Th = sat8i(Th - P.delta);
                    const int mx = t_prev && (l0 < L0);
                    if (mx) { Th = P.rcw + P.alpha; l0 = sat8i(l0 + 1); }
                    const int sub = t_prev && !mx && (l1 < L1);
                    if (sub) { Th = P.rcw + P.alpha - P.delta; l1 = sat8i(l1 + 1); }
                    if (t_prev && !mx && !sub) Th = P.rcw + P.alpha - 2 * P.delta;
                    Th = max(Th, 1);
                    const uint32_t bigm = (P.bf_mode == BF_DTBF || Th >= P.rcw) ? 0xFFFFFFFFu : 0u;  // CDecoder_FAID_2B1C.cpp:6802
                    uint32_t flipped = 0;
                    // regular columns only (weight REGULAR_COL_WEIGHT = 3, :6808): votes + alpha * (hard != hard at BF start)
#define LDPC_BF_DTBF_COL(C)                                                                              \
    if (LDPC_COLW_C##C == 3 && quarter == ((C) & 3)) {                                                   \
        uint32_t x[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, v[4];                                      \
        LDPC_COL_EDGES_C##C(LDPC_BF_VOTE_EDGE)                                                           \
        full_add(x[0], x[1], x[2], v[0], v[1]);                                                          \
        const uint32_t d = diffF[(C) * 8 + word];                                                        \
        const uint32_t da = P.alpha ? d : 0u;                                                            \
        const uint32_t cy = v[0] & da;                                                                   \
        v[0] ^= da;                                                                                      \
        v[2] = v[1] & cy;                                                                                \
        v[1] ^= cy;                                                                                      \
        v[3] = 0;                                                                                        \
        const uint32_t flip = planes4_ge(v, Th);                                                         \
        flipped |= flip;                                                                                 \
        if (flip == 0u) {                                                                                \
            /* nothing to flip in these 32 code bits (the common case): no read-modify-write */          \
        } else if (P.bf_mode == BF_2B1C) {                                                               \
            /* big step: both bits flip; small step: strong bits only lose their second bit (:6805-6813) */ \
            const uint32_t h2 = hard2F[(C) * 8 + word];                                                  \
            const uint32_t fl = (flip & bigm) | (flip & ~h2 & ~bigm);                                    \
            hardF[(C) * 8 + word] ^= fl;                                                                 \
            diffF[(C) * 8 + word] = d ^ fl;                                                              \
            hard2F[(C) * 8 + word] = (bigm & (h2 ^ flip)) | (~bigm & h2 & ~flip);                        \
        } else {                                                                                         \
            hardF[(C) * 8 + word] ^= flip;                                                               \
            diffF[(C) * 8 + word] = d ^ flip;                                                            \
            const uint32_t fl_ = flip;                                                                   \
            if (incr) { LDPC_COL_EDGES_C##C(LDPC_BF_UPD_EDGE) }                                          \
        }                                                                                                \
    }
                    if (quad_active) {
                        LDPC_FOR_EACH_COL(LDPC_BF_DTBF_COL)
                        if (flipped) *flag = 1;
                    }
#undef LDPC_BF_DTBF_COL
                    __syncthreads();
                    t_prev = *flag != 0;
                }
                // the other buffer is written again two barriers from now
                if (quarter == 0 && word == 0) s_flag[(BFiter + 1) & 1][f] = 0;
                if (incr) {  // every vote of this iteration has been read (barrier above): the updated buffer becomes current
                    uint32_t* const tmp = ucur;
                    ucur = uoth;
                    uoth = tmp;
                    const ptrdiff_t d = ucur - uoth;
#pragma unroll
                    for (int a = 0; a < 8; ++a) up[a] += d;
                }
                BFiter++;
            }
        } else
        while (BFiter < P.bf_max_iter) {
            // generic table-driven stage: any regular_col_weight / alpha
            // syndrome of the hard decisions, 96 words (12 layers x 256 rows) per frame
            uint32_t any = 0;
            for (int tau = lane; tau < kUnsatW; tau += 32) {
                const int l = tau >> 3, i = tau & 7;
                uint32_t X = 0;
                for (int e = s_lstart[l]; e < s_lstart[l + 1]; ++e) {
                    const int p = (s_eshift[e] + 32 * i) & 255;
                    const uint32_t* h = hard + s_ecol[e] * 8;
                    X ^= __funnelshift_r(h[p >> 5], h[((p >> 5) + 1) & 7], p & 31);
                }
                unsat[tau] = X;
                any |= X;
            }
            if (!__syncthreads_or(any != 0)) break;  // group-level break (CDecoder_FAID.cpp:6782-6784)

            if (P.bf_mode == BF_PLAIN) {
                // CDecoder_OMSBF.cpp:2994,3327-3335: flip every bit with votes >= min(max(1, max votes), 5)
                uint32_t ge[6] = {0, 0, 0, 0, 0, 0};
                for (int task = lane; task < LDPC_NB * 8; task += 32) {
                    uint32_t v[5];
                    vote_planes(v, unsat, task >> 3, task & 7, s_col_start, s_col_layer, s_col_lshift);
#pragma unroll
                    for (int k = 2; k <= 5; ++k) ge[k] |= planes_ge(v, k);
                }
                int thr = 1;
#pragma unroll
                for (int k = 2; k <= 5; ++k)
                    if (__any_sync(0xFFFFFFFFu, ge[k] != 0)) thr = k;
                for (int task = lane; task < LDPC_NB * 8; task += 32) {
                    uint32_t v[5];
                    vote_planes(v, unsat, task >> 3, task & 7, s_col_start, s_col_layer, s_col_lshift);
                    hard[task] ^= planes_ge(v, thr);
                }
            } else {
                // threshold automaton, per frame (CDecoder_FAID.cpp:6787-6799)
                if (!t_prev) Th = sat8i(Th - P.delta);
                const int mx = t_prev && (l0 < L0);
                if (mx) { Th = P.rcw + P.alpha; l0 = sat8i(l0 + 1); }
                const int sub = t_prev && !mx && (l1 < L1);
                if (sub) { Th = P.rcw + P.alpha - P.delta; l1 = sat8i(l1 + 1); }
                if (t_prev && !mx && !sub) Th = P.rcw + P.alpha - 2 * P.delta;
                Th = max(Th, 1);
                const bool big = Th >= P.rcw;  // CDecoder_FAID_2B1C.cpp:6802
                uint32_t flipped = 0;
                for (int task = lane; task < LDPC_NB * 8; task += 32) {
                    const int c = task >> 3;
                    if (s_colw[c] != P.rcw) continue;  // regular columns only (:6808)
                    uint32_t v[5];
                    vote_planes(v, unsat, c, task & 7, s_col_start, s_col_layer, s_col_lshift);
                    planes_add_const(v, diff[task], P.alpha);
                    const uint32_t flip = planes_ge(v, Th);
                    flipped |= flip;
                    if (P.bf_mode == BF_DTBF || big) {
                        hard[task] ^= flip;
                        diff[task] ^= flip;
                        if (P.bf_mode == BF_2B1C) hard2[task] ^= flip;
                    } else {
                        // small step: strong bits only lose their second bit, weak bits flip (:6812-6813)
                        const uint32_t h2 = hard2[task];
                        const uint32_t fl = flip & ~h2;
                        hard[task] ^= fl;
                        diff[task] ^= fl;
                        hard2[task] = h2 & ~flip;
                    }
                }
                t_prev = __any_sync(0xFFFFFFFFu, flipped != 0);
            }
            __syncwarp();
            BFiter++;
        }
    }
    __syncthreads();

    // ---- outputs ----
    if (lane == 0 && warp == 0) {
        if (P.bf_iters) P.bf_iters[g] = BFiter;
        if (P.its_per_group) P.its_per_group[g] = its;
    }
    if (P.hard_packed) {
        uint32_t* o = P.hard_packed + (size_t)frame * kHW;
        for (int w = lane; w < kHW; w += 32) o[w] = hard[w];
    }
    if (P.decoded) {
        // bit -> byte expansion, 16 code bits per lane per step, 512 B per warp store
        uint4* o = reinterpret_cast<uint4*>(P.decoded + (size_t)frame * kN);
        for (int q = lane; q < kN / 16; q += 32) {
            const uint32_t bits = (hard[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu;
            uint4 v;
            v.x = ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
            v.y = (((bits >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
            v.z = (((bits >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
            v.w = (((bits >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
            o[q] = v;
        }
    }
}

}  // namespace ldpc
