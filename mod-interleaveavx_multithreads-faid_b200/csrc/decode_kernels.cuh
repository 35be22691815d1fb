// decode_kernels.cuh -- layered QC-LDPC message-passing kernels for sm_100a (B200).
//
// Replaces the row loops of CLDPC::Decode (CLDPC.cpp:274-605), CLDPC::Decode_OMS (CDecoder_OMS.cpp:83-745,
// shared verbatim by CDecoder_OMSBF.cpp / CDecoder_OMS_DTBF.cpp), CLDPC::Decode_FAID (CDecoder_FAID.cpp:266-1528)
// and CLDPC::Decode_FAID_2B1C (CDecoder_FAID_2B1C.cpp:183-1340).
//
// Mapping (B200-first, not a translation of the AVX code):
//   * one CTA = 256 threads = the 256 check rows of one block row (layer); the CTA owns a PAIR of frames.
//     Two frames ride in the two 16-bit halves of every 32-bit register (DPX VIMNMX.S16x2 / VIADDMNMX.S16x2 /
//     VIADD.16x2 and VABSDIFF4 are single native instructions on sm_100a; byte-wide SIMD is emulated, see
//     profiles/microbench/).
//   * APP values of the pair live in shared memory, one 32-bit word per code bit (17664 words = 69 KB), stored
//     with a +121 bias so that every quantity of the check-node update is a small unsigned byte.
//     Thread r of a layer touches word col*256 + (shift + r) mod 256 : consecutive lanes -> consecutive banks.
//   * C2V messages never leave the register file: thread r keeps the 4-bit messages of "its" check of every
//     layer (12 layers x 23 edges x 2 frames x 4 bit = 72 registers), fully unrolled so shifts, columns and LUT
//     classes are literals (X-macros of include/ldpc_code_tables.h).  Word k of a layer holds edges 4k..4k+3:
//     frame 0's nibbles in the low half, frame 1's in the high half, so that one shift + one LOP3 unpacks an edge
//     of both frames and ONE multiply-add (FMA pipe) packs it back (see LDPC_P2_TAIL).
//     The words of layers 0..kCvSmemLayers-1 have their home in shared memory (the 44 KB per CTA left beside the
//     APP array at 2 CTAs/SM) and are software-prefetched one layer ahead; the others stay in registers.  This
//     replaces ptxas' local-memory spills (128-register cap), whose reloads stalled on the L1/L2 round trip.
//   * rows of one layer touch disjoint code bits, so a whole layer runs in parallel and is bit-identical to the
//     reference's serial row order; one barrier per layer.
//   * group-of-32 early stop (CDecoder_OMS.cpp:325-327): every CTA publishes "zero syndrome at iteration i"
//     per frame and a packed snapshot of its hard decisions; the group's stop iteration is resolved afterwards
//     by the finalize kernel (bf_kernels.cuh).  A CTA polls the group counter to stop early, which only saves
//     work -- the result does not depend on when (or whether) it observes the stop.
#pragma once
#ifdef LDPC_HOST_EMU
#include "cuda_emu_shim.h"  // tools/emu: the layer arithmetic below compiled for the CPU (test infrastructure)
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "ldpc_code_tables.h"

namespace ldpc {

// KIND_FAID_M / KIND_FAID_EF_M: FAID with V2C LUTs that are monotone and equal for all column-weight classes (true for
// every LUT set the reference ships): the LUT then commutes with the min search and is applied twice per CHECK
// instead of once per EDGE (see LDPC_P1_FAIDM).  KIND_FAID / KIND_FAID_EF stay as the general per-edge path.
// KIND_FAID_ER: KIND_FAID_EF plus the erasure mode (EF_ELIMINATION 2, CDecoder_FAID.cpp:673-680): in the last iterations of a
// frame with few unsatisfied checks, a weight-3 variable node ALL of whose checks are unsatisfied sends v = 0 to the first
// check that visits it.  Needs every row's syndrome bit visible to other threads: one extra KB of shared memory per pair.
enum { KIND_NMS = 0, KIND_OMS = 1, KIND_FAID = 2, KIND_FAID_EF = 3, KIND_FAID_M = 4, KIND_FAID_EF_M = 5, KIND_FAID_ER = 6 };
__host__ __device__ constexpr bool kind_is_faid(int k) { return k >= KIND_FAID; }
__host__ __device__ constexpr bool kind_is_faidm(int k) { return k == KIND_FAID_M || k == KIND_FAID_EF_M; }
__host__ __device__ constexpr bool kind_has_ef(int k) { return k == KIND_FAID_EF || k == KIND_FAID_EF_M || k == KIND_FAID_ER; }

constexpr int kN = LDPC_N, kM = LDPC_M, kK = LDPC_K, kZ = LDPC_Z;
constexpr int kHW = kN / 32;   // packed hard-decision words per frame (552)
constexpr int kMaxIterCap = 1000;  // bound of the API check only; scratch (one 2.2 KB snapshot per frame and iteration) grows with it
constexpr int kThreads = 256;
// Layers whose message words live in shared memory, per kernel kind.  Measured on B200 (1024 groups, 6 iterations,
// gpurun_out/variants_v4.log, variants_v7.log): the register-hungry OMS / FAID kernels gain 8 % / 18 % from 7
// shared-memory layers (no local-memory spills); the lean NMS kernel gains 2 % once two pairs share a CTA.
// At most 7: 2 pairs x (69 KB APP + 6 KB per layer) must fit the 227 KB a CTA may use.
#ifndef LDPC_CV_SMEM_LAYERS_NMS
#define LDPC_CV_SMEM_LAYERS_NMS 7
#endif
#ifndef LDPC_CV_SMEM_LAYERS
#define LDPC_CV_SMEM_LAYERS 7
#endif
__host__ __device__ constexpr int cv_smem_layers(int kind) { return kind == KIND_NMS ? LDPC_CV_SMEM_LAYERS_NMS : LDPC_CV_SMEM_LAYERS; }
// Frame pairs per CTA.  Two pairs (512 threads, 1 CTA/SM) instead of two 256-thread CTAs per SM: the 16 warps of an SM
// then always run the same stretch of the ~100 KB unrolled loop, so they share one instruction stream.  With
// independent CTAs the two streams drift apart as soon as frames converge (snapshot detours) and instruction fetch
// became the top stall (38 % of samples for OMS at 3.6 dB, profiles/r01_oms_v6_ncu_full.md).  Each pair synchronises
// on its own named barrier.
#ifndef LDPC_PAIRS_PER_CTA
#define LDPC_PAIRS_PER_CTA 2
#endif
constexpr int kPairsPerCta = LDPC_PAIRS_PER_CTA;
// word offset, inside a pair's shared memory, of the per-row syndrome words of KIND_FAID_ER (row t: chk0 | chk1 << 16)
__host__ __device__ constexpr int unsat_word_offset(int kind) { return LDPC_N + cv_smem_layers(kind) * 6 * kThreads; }
__host__ __device__ constexpr int pair_smem_words(int kind) { return unsat_word_offset(kind) + (kind == KIND_FAID_ER ? kThreads : 0); }
__host__ __device__ constexpr size_t decode_smem_bytes(int kind) {
    return (size_t)kPairsPerCta * pair_smem_words(kind) * sizeof(uint32_t);
}

// biased representation (see header comment)
// LDPC_FP16_SELECT = 1 (NMS / OMS kinds): every 16-bit quantity carries 0x64 in its high byte, i.e. it IS the fp16
// number 1024 + x.  The is-min select of phase 2 then runs on the FMA pipe, which has slack, instead of the saturated ALU
// pipe: r = sat(|v| - min1) in {0,1} (HADD2.SAT), tp = r * (P2 - P1) + P1 (HFMA2) -- exact on these small integers, and
// valid for any cste_1 / cste_2 order (no MONO requirement).  Integer instructions are unaffected by the constant byte.
// FAID_M kinds use the same two instructions for their threshold select (|v| <= thr).
#ifndef LDPC_FP16_SELECT
#define LDPC_FP16_SELECT 1
#endif
__host__ __device__ constexpr bool kind_fp16(int k) {
    return LDPC_FP16_SELECT && (k == KIND_NMS || k == KIND_OMS || k == KIND_FAID_M || k == KIND_FAID_EF_M);
}
__host__ __device__ constexpr uint32_t hb_of(int k) { return kind_fp16(k) ? 0x64006400u : 0u; }  // high-byte tag per half
constexpr uint32_t kBias = 121;                 // Lb = L + 121            in [90,152]
constexpr uint32_t kBiasM = 24;                 // FAID_M kinds: Yb = L + 24 in [-7,55] (signed 16-bit halves), so that
                                                // v + 31 = relu(min(Yb + (7 - m), 62)) is ONE DPX instruction
__host__ __device__ constexpr int bias_of(int kind) {
    return (kind_is_faidm(kind) ? (int)kBiasM : (int)kBias) + (int)(hb_of(kind) & 0xFFFFu);
}
constexpr uint32_t kU0 = 0x00800080u;           // ub = v + 128
constexpr uint32_t kULo = 0x00610061u;          // v >= -31  <=> ub >= 97
constexpr uint32_t kUHi = 0x009F009Fu;          // v <= +31  <=> ub <= 159
constexpr uint32_t kYLo = 0x00590059u;          // Lb'-1 >= 89
constexpr uint32_t kYHi = 0x00970097u;          // Lb'-1 <= 151
constexpr uint32_t kHardK = 0x7F867F86u;        // Lb + 0x7F86 has bit 15 set  <=>  L > 0
__host__ __device__ constexpr uint32_t hardk_of(int kind) { return (0x7FFFu - (uint32_t)bias_of(kind)) * 0x00010001u; }

// V2C LUTs as PRMT tables: [iteration 1..6][weight class][lo,hi]
struct LutTables {
    uint32_t lut[6][4][2];
    uint32_t lut_ef[6][4][2];
    // FAID_M: thr[x] = largest y with LUT[y] == LUT[x] (127 if that is 7): an edge has t_j == LUT[x] iff |v_j| <= thr[x]
    uint32_t thr[6][2];
    uint32_t thr_ef[6][2];
};

// What defines the transmitted symbols and the channel noise of a frame (frame producer, gen_device.cuh).
struct GenCore {
    const int8_t* output_bits;  // [groups][32*N] two-region layout, or nullptr with `codeword`
    const int8_t* codeword;     // [N] same codeword for every frame (FakeEncoder), or nullptr
    int mod, I;                 // modType, InterleaveModType
    float sigma_d;              // per real dimension: sigma / sqrt(2)
    float scale;                // quantiser scale
    int qbits;                  // quantiser width (4 = float2LimitChar_4bit)
    uint64_t seed, first_frame; // Philox key / global index of the first frame of the launch
    int add_noise;
    int fast_i1;                // I == 1 and 4-byte aligned bit buffers: word-wise bit fetch, no interleaver arithmetic
    // Codeword reuse (CSimulate.cpp:103-117: one Encode() per Run serves 50 noise blocks): group g of the launch takes its
    // transmitted bits from group (group0 + g) / tx_reuse - tx_c0 of output_bits.  tx_reuse <= 1: one codeword group each.
    int tx_reuse;
    uint32_t tx_c0;
    uint64_t group0;
};
__host__ __device__ __forceinline__ int tx_group_of(const GenCore& G, int group) {
    return G.tx_reuse > 1 ? (int)((G.group0 + (uint64_t)group) / (uint64_t)G.tx_reuse - G.tx_c0) : group;
}

struct DecParams {
    const int8_t* llr;        // reference layout: group g at g*32*N; info region then parity region
    const uint8_t* llr_packed;// native layout: frame-major nibbles (used when llr == nullptr)
    GenCore gen;              // fused producer (used when gen_enable): the CTA synthesises its own frames' LLRs
    int gen_enable;
    int8_t* direct_bytes;     // NMS only (no early stop, no BF): decodedBits int8 [group][32][N] written by this kernel,
    uint32_t* direct_packed;  //   or packed decisions [frame][kHW]; finalize_kernel is then not launched at all
    uint32_t* final_hard;     // [frames][planes][kHW]
    uint32_t* snap;           // [frames][max_iter][planes][kHW]
    uint32_t* grp_cnt;        // [groups][max_iter]  frames of the group with zero syndrome at iteration start
    int32_t* first_zero;      // [frames] 1 + index of the first iteration at whose start the syndrome was zero; 0 = never
    int n_frames;
    unsigned int* work_counter;  // [1], zero at launch: next frame pair to hand out (persistent CTAs, see decode_pair_kernel)
    int max_iter;
    int planes;               // 1, or 2 when the 2B1C second bit (|L| >= hard2_thr) is needed
    int hard2_thr;
    int puncture_tail;
    int factor_1, factor_2;   // NMS
    int nms_fast;             // both factors in [0, 2114]: (min * factor) cannot wrap 16 bits, the scaling runs on both halves at once
    unsigned long long* dbg;  // LDPC_DEBUG_BOUNDS builds: [0] violation count, [1] first (code << 32 | value); else unused
    int exp_fault;            // LDPC_DEBUG_BOUNDS builds, LDPC_B200_DEBUG_FAULT: evaluate one out-of-range APP offset (negative control)
    int exp_noload;           // experiment switch (LDPC_B200_EXP_NOLOAD): skip the LLR load (timing of the load phase only)
    int no_skew;              // experiment switch (LDPC_B200_NO_SKEW): both halves of a CTA start their first item together
    uint32_t oms_norm[2], oms_boost[2];  // 8-entry byte LUTs: 64 + cste as a function of the (clipped) minimum
    int oms_floor_err, oms_floor_iter;
    int ef_floor_err, ef_floor_iter;
    int err_sat;              // 255 (OMS family, unsigned saturation) or 127 (FAID family, signed)
    LutTables luts;           // FAID kinds: per handle, in the kernel's parameter (constant) bank -- two handles with different
                              // LUT sets can share a device
};


// device copy of the QC description (filled from include/ldpc_code_tables.h by the host at create())
struct CodeTables {
    uint8_t circ_col[LDPC_NCIRC], circ_shift[LDPC_NCIRC], col_layer[LDPC_NCIRC], col_lshift[LDPC_NCIRC];
    uint16_t layer_start[LDPC_MB + 1], col_start[LDPC_NB + 1];
    uint8_t col_weight[LDPC_NB];
    uint32_t hpinv[LDPC_MB][LDPC_MB][LDPC_Z / 32];
};

struct IterCtx {
    uint32_t lut[4][2];
    uint32_t lut_ef[4][2];
    uint32_t thr[2], thr_ef[2];
    uint32_t chk0, chk1;      // bit L: row (layer L, r) of frame 0/1 unsatisfied at iteration start
    uint32_t lane_ok;         // 16x2 mask: frame's error_sum below the floor count
    int special_active;       // remaining iterations <= floor_iter_thresh
};

// prmt.b32 in its generic form: selector bit 3 replicates the sign (msb) of the selected byte over the whole
// output byte.  (__byte_perm masks every selector nibble to 3 bits, so it cannot express this.)
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t sel) {
#ifdef LDPC_HOST_EMU
    return emu_prmt(a, 0u, sel);
#else
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
    return d;
#endif
}
__device__ __forceinline__ uint32_t sel32(uint32_t m, uint32_t a, uint32_t b) { return (m & a) | (~m & b); }
#if LDPC_DEBUG_BOUNDS && !defined(LDPC_HOST_EMU)
// dbg[0] = number of violations, dbg[1] = (code << 32) | value of the first one
enum { DBG_APP = 1, DBG_CV = 2, DBG_SNAP = 3, DBG_HARD = 4, DBG_LLR = 5, DBG_UNSAT = 6, DBG_FIN = 7 };
static __device__ __noinline__ void ldpc_bounds_fail(unsigned long long* dbg, int code, unsigned val) {
    if (!dbg) return;
    atomicAdd(&dbg[0], 1ull);
    atomicCAS(&dbg[1], 0ull, ((unsigned long long)code << 32) | val);
}
#define LDPC_CHECK(dbg, cond, code, val) do { if (!(cond)) ldpc_bounds_fail((dbg), (code), (unsigned)(val)); } while (0)
// byte offset of an APP word relative to the CTA's shared memory: inside this pair's APP array, word aligned
__device__ __forceinline__ uint32_t* ldpc_app_checked(uint32_t* app, uint32_t byte_off, uint32_t pbase, unsigned long long* dbg) {
    LDPC_CHECK(dbg, byte_off >= pbase && byte_off < pbase + LDPC_N * 4u && (byte_off & 3u) == 0u, DBG_APP, byte_off);
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(app) + byte_off);
}
#else
#define LDPC_CHECK(dbg, cond, code, val) do { } while (0)
#endif
// fp16x2 arithmetic on raw bit patterns (exact here: all operands are integers below 2048)
#ifdef LDPC_HOST_EMU
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { return emu_h2_sub(a, b, false); }
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { return emu_h2_add(a, b); }
__device__ __forceinline__ uint32_t h2_sub_sat(uint32_t a, uint32_t b) { return emu_h2_sub(a, b, true); }
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { return emu_h2_fma(a, b, c); }
__device__ __forceinline__ uint32_t h2_abs(uint32_t a) { return a & 0x7FFF7FFFu; }
#else
__device__ __forceinline__ __half2 as_h2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }
__device__ __forceinline__ uint32_t as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) { return as_u32(__hsub2(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { return as_u32(__hadd2(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t h2_sub_sat(uint32_t a, uint32_t b) { return as_u32(__hsub2_sat(as_h2(a), as_h2(b))); }
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { return as_u32(__hfma2(as_h2(a), as_h2(b), as_h2(c))); }
__device__ __forceinline__ uint32_t h2_abs(uint32_t a) { return as_u32(__habs2(as_h2(a))); }  // folds into a source modifier
#endif
__device__ __forceinline__ uint32_t expand2(uint32_t b0, uint32_t b1) {  // two booleans -> 16x2 mask
    return (b0 ? 0x0000FFFFu : 0u) | (b1 ? 0xFFFF0000u : 0u);
}
// 8-entry byte LUT lookup of both halves (index in the low byte of each half, 0..7)
__device__ __forceinline__ uint32_t lut8(uint32_t lo, uint32_t hi, uint32_t idx) {
    return __byte_perm(lo, hi, __byte_perm(idx, 0, 0x4420)) & 0x00FF00FFu;
}
// CLDPC.cpp:342-363 on one half: ((min * factor) mod 2^16) >> 5, saturating pack, min with 7
__device__ __forceinline__ uint32_t nms_scale1(uint32_t m, int factor) {
    uint32_t p = ((m & 0xFFu) * (uint32_t)(factor & 0xFFFF)) & 0xFFFFu;
    p >>= 5;
    return p < 7u ? p : 7u;
}
__device__ __forceinline__ uint32_t nms_scale(uint32_t m2, int factor) {
    return nms_scale1(m2 & 0xFFFFu, factor) | (nms_scale1(m2 >> 16, factor) << 16);
}

// The same for both halves at once, valid when 31 * factor < 2^16 (no 16-bit wrap, host check -> DecParams.nms_fast):
// untag, one multiply, one shift, one mask, one unsigned min.
__device__ __forceinline__ uint32_t nms_scale16(uint32_t m2, int factor) {
    const uint32_t p = (m2 & 0x00FF00FFu) * (uint32_t)factor;
#ifdef LDPC_HOST_EMU
    const uint32_t q = (p >> 5) & 0x07FF07FFu;
    return emu_mk(emu_lo(q) < 7 ? emu_lo(q) : 7, emu_hi(q) < 7 ? emu_hi(q) : 7);
#else
    return __vminu2((p >> 5) & 0x07FF07FFu, 0x00070007u);
#endif
}

// Message nibbles (m + 8) of edge j of both frames, moved down to bits 0..3 of each half (upper bits: don't care).
// (Doing the shift as the high half of a multiply, i.e. on the FMA pipe, was measured: IMAD.HI issues at half rate.)
#define LDPC_NIB(j) (((j) & 3) == 0 ? cv[(j) >> 2] : (cv[(j) >> 2] >> (4 * ((j) & 3))))
// Repacking is arithmetic: word = sum_i (cmo_i - 56) * 16^i per half, accumulated mod 2^32 by one IMAD per edge
// (the true value of each half fits 16 bits, so intermediate carries across the halves cancel).  First edge of a
// word adds the constant -56 * (16^0 + .. + 16^(n-1)) * 0x10001 for the n edges the word holds.
#define LDPC_PACK_N(j) ((j) + 4 <= DEG ? 4 : DEG - (j))
#define LDPC_PACK_INIT(j) (0u - (56u + (HB & 0xFFFFu)) * 0x00010001u * (LDPC_PACK_N(j) == 4 ? 0x1111u : LDPC_PACK_N(j) == 3 ? 0x111u : LDPC_PACK_N(j) == 2 ? 0x11u : 0x1u))

// Byte offset of check row r's word inside a 256-word block column: ((r + shift) mod 256) * 4.
// 69 of the 275 circulants have shift 0.  pbase = byte offset of this pair's APP array inside the CTA's shared memory
// (a multiple of 1024) and rr = 4 * row + pbase; the LOP3 that wraps the row index also re-inserts the base:
// ((rr + 4 s) & 1020) | pbase.
#define LDPC_OFF(s) ((s) == 0 ? rr : (((rr + 4u * (s)) & 1020u) | pbase))
#if LDPC_DEBUG_BOUNDS && !defined(LDPC_HOST_EMU)
#define LDPC_APP(c, off) (*ldpc_app_checked(app, (uint32_t)((c) * 1024) + (off), pbase, P.dbg))
#else
#define LDPC_APP(c, off) (*reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(app) + (c) * 1024 + (off)))
#endif

constexpr uint32_t kP0 = 0x00400040u;    // selected constants are kept as 64 +- c (low byte of each half)
constexpr uint32_t kNeg = 0x00800080u;   // |x - 128| = 64 - c : VABSDIFF4 against this flag negates
constexpr uint32_t kEC = 0x08000800u;    // e = 2048 - 8 a
constexpr uint32_t kSumLo = 0x00A100A1u; // 161 <= ub + (64 +- c) <= 223  <=>  -31 <= L' <= 31
constexpr uint32_t kSumHi = 0x00DF00DFu;
constexpr uint32_t kSumBias = 0xFF5FFF5Fu;  // - 161

#ifndef LDPC_MIN2_TREE
#define LDPC_MIN2_TREE 1
#endif
#ifndef LDPC_NMS_SCALE16
#define LDPC_NMS_SCALE16 1
#endif
// -DLDPC_DEBUG_BOUNDS=1: every shared-memory access of the layer code (APP words, message words), every index into the
// snapshot / hard-decision / LLR buffers is range-checked on the device; violations are counted in DecParams.dbg and read
// back with ldpc_b200_debug_bounds().  The GPU pool keeps compute-sanitizer closed; this build is its substitute and is run
// by tests/test_gpu_bounds_debug.py on all six DecodeMethods.  Off (and free) in the shipped library.
#ifndef LDPC_DEBUG_BOUNDS
#define LDPC_DEBUG_BOUNDS 0
#endif
// experiments: which pipe some 16x2 additions of phase 2 / the magnitude of phase 1 run on
// 0 (default): one CTA per two frame pairs, both pairs start together and -- doing identical work -- stay within a layer of
// each other, so the SM's 16 warps share ONE instruction stream.
// 1: persistent CTAs whose two halves take frame pairs from a ticket counter and run half an item apart, so that a half that
// waits on HBM (LLR load, result store) overlaps with the other half's arithmetic.  Measured on B200 (same box,
// profiles/r02_persistent_ab.md): 12 % SLOWER for NMS (6.64 vs 5.93 ms) -- the unrolled iteration is 91 KB of code against a
// 32 KB L1.5 instruction cache, two halves at different layers are two instruction streams from L2, and `no_instruction`
// goes from 5.7 % to 27.9 % of all warp-stall samples, far more than the 8 % of load-phase stalls it hides.
#ifndef LDPC_PERSISTENT
#define LDPC_PERSISTENT 0
#endif
// Whether a layer's message words have a shared-memory home (bit 0) / a prefetch for the next layer (bit 1) is a template flag
// of the layer function instead of a test of the pointer (which ptxas cannot prove non-null: a predicate per layer, the dead
// path's register moves, 240 B of spills in the NMS kernel instead of 128).  Which flags pay differs per kernel family -- ptxas'
// register allocation of the 128-register layer bodies is not monotone in what it is told.  Measured on B200, same box, ms per
// 1024 groups at 3.6 dB (profiles/r02_nms_ab_exp15.log, r02_nms_ab_exp17.log):
//   NMS     none 5.76   home 5.66   prefetch 5.38   both 5.27  (83.0 -> 90.7 Gbit/s)
//   OMS     none 7.17   home 6.95   prefetch 6.96   both 7.02
//   FAID_M  none 10.73  home 10.53  prefetch 10.55  both 11.23 (spills 80 -> 690 B)
//   FAID_EF_M none 11.11 home 11.12 prefetch 11.49  both 11.92
// LDPC_CV_STATIC=0 switches all of them off.
#ifndef LDPC_CV_STATIC
#define LDPC_CV_STATIC 1
#endif
#ifndef LDPC_CV_STATIC_NMS
#define LDPC_CV_STATIC_NMS 3
#endif
#ifndef LDPC_CV_STATIC_OMS
#define LDPC_CV_STATIC_OMS 1
#endif
#ifndef LDPC_CV_STATIC_FAID
#define LDPC_CV_STATIC_FAID 1   // the monotone-LUT kinds; the general per-edge-LUT kinds keep the pointer tests (not measured)
#endif
#define LDPC_CV_STATIC_BITS (KIND == KIND_NMS ? LDPC_CV_STATIC_NMS : KIND == KIND_OMS ? LDPC_CV_STATIC_OMS : kind_is_faidm(KIND) ? LDPC_CV_STATIC_FAID : 0)
#define LDPC_CV_STATIC_HOME_K (LDPC_CV_STATIC && (LDPC_CV_STATIC_BITS & 1))
#define LDPC_CV_STATIC_PRE_K (LDPC_CV_STATIC && (LDPC_CV_STATIC_BITS & 2))
// experiment: prefetch of the next layer's message words after phase 2 (6 registers fewer live during it) instead of between the phases
#ifndef LDPC_PRE_LATE
#define LDPC_PRE_LATE 0
#endif
#define LDPC_PRE_LATE_K (LDPC_PRE_LATE && (KIND == KIND_NMS || KIND == KIND_OMS))
#ifndef LDPC_P2_ADD_ALU
#define LDPC_P2_ADD_ALU 0
#endif
// KIND_FAID_M: 1 = the word (v + 31 | sign << 15) of an edge waits for phase 2 in a register and |v| is re-derived there on the
// FMA pipe (two HADD2); 0 = it is parked in the APP array (STS in phase 1, LDS in phase 2) and |v| stays in the register.
// Measured on B200, same box (profiles/r02_nms_ab_exp9.log): FAID3 + DTBF 11.05 -> 10.68 ms per 1024 groups at 3.6 dB, 9.42 ->
// 9.07 at 4.2 dB.  The error-floor kind (KIND_FAID_EF_M, hybrid decoder) spills with it (44 -> 388 B) and gets 4 % slower, so
// it keeps the shared-memory form.
#ifndef LDPC_FAIDM_REG
#define LDPC_FAIDM_REG 1
#endif
#define LDPC_FAIDM_IN_REG (LDPC_FAIDM_REG && HB && KIND == KIND_FAID_M)
#ifndef LDPC_ABS_FP16
#define LDPC_ABS_FP16 0
#endif
#if LDPC_MIN2_TREE
// Two smallest of the DEG magnitudes of a check (min2 == min1 on ties, CLDPC.h:68) as a tournament of TRIPLES.
//   triple (x0, x1, x2):  m = min3, M = max3 (two ALU-pipe instructions),
//                         med = x2 + (x0 - m) - (M - x1)  (four HADD2 on the FMA pipe, which has slack; see min2_triple)
//   the smallest value of a set is the smallest triple minimum; the second smallest is
//   min(second smallest of the triple minima, smallest median)  -- every median is an element other than the minimum,
//   and if the runner-up is not itself a triple minimum it is the median of its triple.
// Triple minima recurse through three levels (23 -> 9 -> 3 -> 1, 22 -> 8 -> 4 -> 2 -> 1); medians are folded into one
// running value with min3.  27 ALU-pipe instructions per 23-edge check instead of the 57 of a running (min1, min2) pair
// fed two values at a time.  Everything is resolved at compile time from the literal edge index.
struct Min2Tree {
    uint32_t a0, a1;   // level 1: pending values of the current triple
    uint32_t b0, b1;   // level 2
    uint32_t c[4];     // level 3 inputs (at most 4)
    uint32_t h;        // smallest median so far (starts at the reference's initial min2 = 31)
    uint32_t ha, hb;   // one pending median per level, so that two fold into h with a single min3
};
// experiment: which triples take the XOR median (2 ALU-pipe instructions) instead of the fp16 one (4 on the FMA pipe):
// bit 0 = even level-1 triples, bit 1 = odd level-1 triples, bit 2 = level 2, bit 3 = level 3
#ifndef LDPC_MED_XOR_MASK
#define LDPC_MED_XOR_MASK 0
#endif
template <bool FP16_>
__device__ __forceinline__ void min2_triple(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t& m, uint32_t& med) {
    constexpr bool FP16 = FP16_;
    m = __vimin3_s16x2(x0, x1, x2);
    const uint32_t M = __vimax3_s16x2(x0, x1, x2);
    // FP16 (values carry the 0x64 tag = fp16 1024 + x): x0 - m and M - x1 are exact small fp16 numbers, adding them to
    // a tagged value gives the tagged result: four HADD2.  Untagged values: the multiset {x0,x1,x2} = {m,med,M}, so
    // med = x0 ^ x1 ^ x2 ^ m ^ M (two LOP3; a 16x2 subtraction would cost three instructions).
    if (FP16) med = h2_sub(h2_add(x2, h2_sub(x0, m)), h2_sub(M, x1));
    else med = (x0 ^ x1 ^ x2) ^ (m ^ M);
}
template <int DEG>
struct Min2Shape {
    static constexpr int T1 = DEG / 3, N2 = T1 + DEG % 3, T2 = N2 / 3, N3 = T2 + N2 % 3;
    static_assert(N3 >= 2 && N3 <= 4, "check degree outside the range this tournament was laid out for");
};
template <int DEG, bool FP16, int I>
__device__ __forceinline__ void min2_feed3(Min2Tree& s, uint32_t z) { s.c[I] = z; }
template <int DEG, bool FP16, int I>
__device__ __forceinline__ void min2_feed2(Min2Tree& s, uint32_t y) {
    using Sh = Min2Shape<DEG>;
    if constexpr (I < 3 * Sh::T2) {
        if constexpr (I % 3 == 0) s.b0 = y;
        else if constexpr (I % 3 == 1) s.b1 = y;
        else {
            uint32_t m, med;
            min2_triple<(FP16 && !(LDPC_MED_XOR_MASK & 4))>(s.b0, s.b1, y, m, med);
            if constexpr ((I / 3) % 2 == 0) s.hb = med; else s.h = __vimin3_s16x2(s.h, s.hb, med);
            min2_feed3<DEG, FP16, I / 3>(s, m);
        }
    } else {
        min2_feed3<DEG, FP16, Sh::T2 + (I - 3 * Sh::T2)>(s, y);
    }
}
template <int DEG, bool FP16, int J>
__device__ __forceinline__ void min2_feed(Min2Tree& s, uint32_t x) {
    using Sh = Min2Shape<DEG>;
    if constexpr (J < 3 * Sh::T1) {
        if constexpr (J % 3 == 0) s.a0 = x;
        else if constexpr (J % 3 == 1) s.a1 = x;
        else {
            uint32_t m, med;
            min2_triple<(FP16 && !(LDPC_MED_XOR_MASK & (((J / 3) % 2 == 0) ? 1 : 2)))>(s.a0, s.a1, x, m, med);
            if constexpr ((J / 3) % 2 == 0) s.ha = med; else s.h = __vimin3_s16x2(s.h, s.ha, med);
            min2_feed2<DEG, FP16, J / 3>(s, m);
        }
    } else {
        min2_feed2<DEG, FP16, Sh::T1 + (J - 3 * Sh::T1)>(s, x);
    }
}
// after the last edge: resolve level 3 and the pending medians.  cap = the reference's initial value of both minima.
template <int DEG, bool FP16>
__device__ __forceinline__ void min2_finish(Min2Tree& s, uint32_t cap, bool cap_min1, uint32_t& min1, uint32_t& min2) {
    using Sh = Min2Shape<DEG>;
    uint32_t h = s.h;
    // pending medians: level 1 has T1 of them, level 2 has T2; an odd count leaves one unfolded
    if constexpr (Sh::T1 % 2 == 1 && Sh::T2 % 2 == 1) h = __vimin3_s16x2(h, s.ha, s.hb);
    else if constexpr (Sh::T1 % 2 == 1) h = __vmins2(h, s.ha);
    else if constexpr (Sh::T2 % 2 == 1) h = __vmins2(h, s.hb);
    uint32_t m, r;  // smallest / second smallest of the level-3 inputs
    if constexpr (Sh::N3 == 2) {
        m = __vmins2(s.c[0], s.c[1]);
        r = __vmaxs2(s.c[0], s.c[1]);
    } else if constexpr (Sh::N3 == 3) {
        min2_triple<(FP16 && !(LDPC_MED_XOR_MASK & 8))>(s.c[0], s.c[1], s.c[2], m, r);
    } else {
        uint32_t m3, med;
        min2_triple<(FP16 && !(LDPC_MED_XOR_MASK & 8))>(s.c[0], s.c[1], s.c[2], m3, med);
        m = __vmins2(m3, s.c[3]);
        r = __vmins2(med, __vmaxs2(m3, s.c[3]));
    }
    min1 = cap_min1 ? __vmins2(m, cap) : m;
    min2 = __vmins2(h, r);
    (void)cap;
}
#define LDPC_MIN2_FEED(j, x) min2_feed<DEG, kind_fp16(KIND), j>(mt, x);
#else
// running two smallest values, fed two candidates at a time (5 instructions per 2 edges)
#define LDPC_MIN2_PAIR(x0, x1)                                    \
    {                                                             \
        const uint32_t lo = __vmins2(x0, x1), hi = __vmaxs2(x0, x1); \
        min2 = __vimin3_s16x2(min2, hi, __vmaxs2(min1, lo));      \
        min1 = __vmins2(min1, lo);                                \
    }
#define LDPC_MIN2_ONE(x0)                                         \
    {                                                             \
        min2 = __vmins2(min2, __vmaxs2(min1, x0));                \
        min1 = __vmins2(min1, x0);                                \
    }
#define LDPC_MIN2_FEED(j, x)                                      \
    if (((j) & 1) == 0) {                                         \
        if ((j) == DEG - 1) LDPC_MIN2_ONE(x) else held = (x);     \
    } else LDPC_MIN2_PAIR(held, x)
#endif

// ---- phase 1: V2C, sign parity, two smallest magnitudes -------------------------------------------------
#define LDPC_P1_COMMON(j, c, s)                                                  \
    const uint32_t off = LDPC_OFF(s);                                            \
    const uint32_t Lb = LDPC_APP(c, off);                                        \
    const uint32_t nibc = ~LDPC_NIB(j) & 0x000F000Fu; /* 7 - m per half */

#define LDPC_P1_MS(j, c, s, w)                                                   \
    {                                                                            \
        LDPC_P1_COMMON(j, c, s)                                                  \
        const uint32_t u = __viaddmax_s16x2(Lb, nibc, kULo + HB);                \
        ub[j] = u;                                                               \
        if (((j) & 1) == 0) { if ((j) == DEG - 1) S ^= u; else uheld = u; }      \
        else S = S ^ uheld ^ u;                                                  \
        const uint32_t a = (LDPC_ABS_FP16 && HB) ? h2_add(h2_abs(h2_sub(u, 0x64806480u)), 0x64006400u) : __vabsdiffu4(u, kU0); \
        LDPC_MIN2_FEED(j, a)                                                     \
    }

// Erasure (KIND_FAID_ER): block column c has weight 3 and this layer is its first one.  The variable node of row t sits at
// word (t + s) & 255; its other two checks are rows (t + s - s1) & 255 of layer l1 and (t + s - s2) & 255 of layer l2.
// Erased halves: v = 0 (u = 128) and, FAID2_SIGN_BACKTRACK with v = 0, the sign of L itself (bit 15 of Lb + 0x8000 - 121 <=> L >= 0).
#define LDPC_UNSAT(d) (*reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(app) + 4 * unsat_word_offset(KIND) + ((((rr) + 4u * ((uint32_t)(d) & 255u)) & 1020u) | pbase)))
#define LDPC_P1_FAID(j, c, s, w)                                                 \
    {                                                                            \
        LDPC_P1_COMMON(j, c, s)                                                  \
        uint32_t u = __vmins2(__viaddmax_s16x2(Lb, nibc, kULo), kUHi);           \
        uint32_t w2 = u * 256u - ((nibc & 0x00080008u) << 4);                    \
        if (KIND == KIND_FAID_ER && LDPC_COLW_C##c == 3 && LDPC_COL3_L0_C##c == kLY) {                   \
            const uint32_t ua = LDPC_UNSAT((s) - LDPC_COL3_S1_C##c) >> (LDPC_COL3_L1_C##c & 15);         \
            const uint32_t uc = LDPC_UNSAT((s) - LDPC_COL3_S2_C##c) >> (LDPC_COL3_L2_C##c & 15);         \
            const uint32_t er = eef & ((ua & uc & 0x00010001u) * 0xFFFFu);                               \
            u = sel32(er, kU0, u);                                                                       \
            w2 = sel32(er, Lb + (0x8000u - kBias) * 0x00010001u, w2); /* bit 15 <=> L >= 0 */              \
        }                                                                                                \
        if (((j) & 1) == 0) { if ((j) == DEG - 1) S ^= w2; else uheld = w2; }    \
        else S = S ^ uheld ^ w2;                                                 \
        LDPC_APP(c, off) = (w2 & 0x80008000u) | u;                               \
        const uint32_t a7 = __vmins2(__vabsdiffu4(u, kU0), 0x00070007u);         \
        uint32_t t = lut8(cx.lut[w][0], cx.lut[w][1], a7);                       \
        if (kind_has_ef(KIND)) t = sel32(eef, lut8(cx.lut_ef[w][0], cx.lut_ef[w][1], a7), t); \
        ub[j] = t;                                                               \
        LDPC_MIN2_FEED(j, t)                                                     \
    }

// FAID with monotone LUTs.  APP words carry Yb = L + 24, so up = v + 31 (clamped to [0,62]) is one instruction.
// Sign with FAID2_SIGN_BACKTRACK (CDecoder_FAID.cpp:681-682: sign of L when v == 0, and then L = m): v, m >= 0 <=>
// 16 v + m >= 0 <=> bit 15 of W = 16 (16 (up - 31) + m) + 32768 = 256 up - 16 (7 - m) + 24944: two IMADs.
// Between the phases the APP word holds up | nonneg << 15; ub[j] keeps |v| for the threshold select.
#define LDPC_P1_FAIDM(j, c, s, w)                                                \
    {                                                                            \
        LDPC_P1_COMMON(j, c, s)                                                  \
        /* fp16 form: Lb carries the 0x64 tag, nibcF = (7 - m) - 0x6400 removes it in the same add (the OR merges into  \
           the unpacking LOP3); the sign arithmetic absorbs the constant: 0x61706170 + 16 * 0x9C009C00 mod 2^32 */    \
        const uint32_t nibcF = HB ? (nibc | 0x9C009C00u) : nibc;                 \
        const uint32_t up = __viaddmin_s16x2_relu(Lb, nibcF, 0x003E003Eu);       \
        const uint32_t W = (up * 256u + (HB ? 0x217A2170u : 0x61706170u)) - nibcF * 16u; \
        if (((j) & 1) == 0) { if ((j) == DEG - 1) S ^= W; else uheld = W; }      \
        else S = S ^ uheld ^ W;                                                  \
        const uint32_t a = __vabsdiffu4(up, 0x001F001Fu + HB); /* |v|, tagged by the |0 - 0x64| of the high byte */ \
        if (LDPC_FAIDM_IN_REG) ub[j] = (W & 0x80008000u) | up;                   \
        else { LDPC_APP(c, off) = (W & 0x80008000u) | up; ub[j] = a; }           \
        LDPC_MIN2_FEED(j, a)                                                     \
    }

// is-min <=> |v| <= thr: integer form tp = max(P1 - 8 relu(|v| - thr), P2) (nthr = -thr per half);
// fp16 form r = sat(|v| - thr) in {0,1}, tp = r * (P2 - P1) + P1 on the FMA pipe
#define LDPC_P2_FAIDM(j, c, s, w)                                                 \
    {                                                                             \
        const uint32_t off = LDPC_OFF(s);                                         \
        const uint32_t q = LDPC_FAIDM_IN_REG ? ub[j] : LDPC_APP(c, off);          \
        const uint32_t up = q & 0x003F003Fu;                                      \
        uint32_t tp;                                                              \
        if (LDPC_FAIDM_IN_REG) /* |v| = |(1024 + up) - (1024 + 31)|, r = sat(|v| - thr), all exact in fp16 */ \
            tp = h2_fma(h2_sub_sat(h2_abs(h2_sub(up | 0x64006400u, 0x641F641Fu)), nthr), Dh, P1c); \
        else if (HB) tp = h2_fma(h2_sub_sat(ub[j], nthr), Dh, P1c);               \
        else tp = __viaddmax_s16x2(P1big - __viaddmax_s16x2_relu(ub[j], nthr, 0u) * 8u, 0xF800F800u, P2c); \
        const uint32_t fl = ((q >> 8) ^ Sp) & kNeg;                               \
        const uint32_t cmo = __vabsdiffu4(tp, fl); /* 64 + c or 64 - c (+ tag) */ \
        const uint32_t y = __viaddmin_s16x2_relu(up, __vadd2(cmo, 0xFFC0FFC0u - HB), 0x003E003Eu); /* L' + 31 */ \
        LDPC_APP(c, off) = __vadd2(y, HB ? 0x63F963F9u : 0xFFF9FFF9u); /* - 7 (+ tag), per half */ \
        nw = ((j) & 3) == 0 ? cmo + LDPC_PACK_INIT(j) : cmo * (1u << (4 * ((j) & 3))) + nw; \
        if (((j) & 3) == 3 || (j) == DEG - 1) {                                   \
            if (LDPC_CV_STATIC_HOME_K ? HOME : (cv_home != nullptr)) { LDPC_CV_CHECK(&cv_home[((j) >> 2) * kThreads]) cv_home[((j) >> 2) * kThreads] = nw; } else cv[(j) >> 2] = nw;  \
        }                                                                         \
    }

// message words live right behind the pair's APP array: [kN, kN + layers * 6 * 256) words of the pair's region
#define LDPC_CV_CHECK(ptr)                                                                                              \
    LDPC_CHECK(P.dbg, (uint32_t)(reinterpret_cast<const char*>(ptr) - reinterpret_cast<const char*>(app)) >= pbase + LDPC_N * 4u && \
                      (uint32_t)(reinterpret_cast<const char*>(ptr) - reinterpret_cast<const char*>(app)) < pbase + (LDPC_N + cv_smem_layers(KIND) * 6 * kThreads) * 4u, \
               DBG_CV, (uint32_t)(reinterpret_cast<const char*>(ptr) - reinterpret_cast<const char*>(app)));

// ---- phase 2: C2V select, sign, APP write-back, message repack ---------------------------------------------
// tp = 64 + (is-min ? c1 : c2).  MONO (c1 >= c2 for every reachable pair of minima, checked on the host):
//   tp = max(P1 - 8 (a - min1), P2) in one VIADDMNMX; otherwise mask-select.
#define LDPC_P2_SELECT(a_)                                                                         \
    uint32_t tp;                                                                                   \
    if (kind_fp16(KIND)) tp = h2_fma(h2_sub_sat(a_, min1), Dh, P1c);                               \
    else if (MONO) tp = __viaddmax_s16x2(kEC - (a_) * 8u, Qp, P2c);                                 \
    else tp = sel32(__viaddmin_s16x2(a_, nmin1, 0x00010001u) * 0xFFFFu, P2c, P1c);

#define LDPC_P2_TAIL(j, c)                                                        \
    const uint32_t cmo = __vabsdiffu4(tp, fl); /* 64 + c or 64 - c */             \
    /* one DPX op clamps to [0, 62] = L' + 31:  max(min((u - 161) + cmo, 62), 0) */ \
    const uint32_t y = __viaddmin_s16x2_relu(__vadd2(u, kSumBias - 2u * HB), cmo, 0x003E003Eu); \
    LDPC_APP(c, off) = LDPC_P2_ADD_ALU ? __viaddmax_s16x2(y, 0x005A005Au + HB, 0u) : __vadd2(y, 0x005A005Au + HB); \
    nw = ((j) & 3) == 0 ? cmo + LDPC_PACK_INIT(j) : cmo * (1u << (4 * ((j) & 3))) + nw; \
    if (((j) & 3) == 3 || (j) == DEG - 1) {                                       \
        if (LDPC_CV_STATIC_HOME_K ? HOME : (cv_home != nullptr)) { LDPC_CV_CHECK(&cv_home[((j) >> 2) * kThreads]) cv_home[((j) >> 2) * kThreads] = nw; } else cv[(j) >> 2] = nw;  \
    }

#define LDPC_P2_MS(j, c, s, w)                                                    \
    {                                                                             \
        const uint32_t off = LDPC_OFF(s);                                         \
        const uint32_t u = ub[j];                                                 \
        const uint32_t a = __vabsdiffu4(u, kU0);  /* ptxas keeps phase 1's value when registers allow */ \
        LDPC_P2_SELECT(a)                                                         \
        const uint32_t fl = (Sp ^ u) & kNeg;                                      \
        LDPC_P2_TAIL(j, c)                                                        \
    }

#define LDPC_P2_FAID(j, c, s, w)                                                  \
    {                                                                             \
        const uint32_t off = LDPC_OFF(s);                                         \
        const uint32_t q = LDPC_APP(c, off);                                      \
        const uint32_t u = q & 0x00FF00FFu;                                       \
        LDPC_P2_SELECT(ub[j])                                                     \
        const uint32_t fl = ((Sp ^ q) >> 8) & kNeg;                               \
        LDPC_P2_TAIL(j, c)                                                        \
    }

// One layer.  cv[6] = this thread's packed messages of the layer.  cv_home: shared-memory home of those words
// (updated words are stored there; nullptr = the words live in cv itself).  pre / cv_next: prefetch of the NEXT
// layer's words from shared memory, issued between the two phases (nullptr = next layer is register-resident).
#if LDPC_MIN2_TREE
// |v| of the min-sum kinds is not clamped at +31 (CLDPC.cpp:330), so their min1 needs the reference's initial value as a cap
#define LDPC_MIN2_DECL Min2Tree mt; mt.h = min2;
#define LDPC_MIN2_FINISH min2_finish<DEG, kind_fp16(KIND)>(mt, 0x001F001Fu + HB, KIND == KIND_NMS || KIND == KIND_OMS, min1, min2);
#else
#define LDPC_MIN2_DECL
#define LDPC_MIN2_FINISH
#endif
#define LDPC_DEF_LAYER(LY)                                                                              \
    template <int KIND, bool MONO, bool HOME, bool PRE>                                                 \
    __device__ __forceinline__ void layer_##LY(uint32_t* __restrict__ app, const uint32_t rr, const uint32_t pbase, \
                                               uint32_t (&cv)[6],                                       \
                                               uint32_t* cv_home, const uint32_t* pre, uint32_t (&cv_next)[6], \
                                               const IterCtx& cx, const DecParams& P) {                 \
        constexpr int DEG = LDPC_DEG_L##LY;                                                             \
        constexpr uint32_t HB = hb_of(KIND);                                                            \
        constexpr int kLY = LY;                                                                         \
        (void)pbase; (void)kLY;                                                                         \
        uint32_t ub[LDPC_MAXDEG];                                                                       \
        uint32_t S = 0, min1 = 0x001F001Fu + HB, min2 = 0x001F001Fu + HB, held = 0, uheld = 0;           \
        const uint32_t rowsel = expand2((cx.chk0 >> LY) & 1u, (cx.chk1 >> LY) & 1u) & cx.lane_ok;       \
        const uint32_t eef = cx.special_active ? rowsel : 0u;                                           \
        (void)eef; (void)held; (void)uheld;                                                             \
        LDPC_MIN2_DECL                                                                                  \
        if (KIND == KIND_NMS || KIND == KIND_OMS) {                                                     \
            LDPC_EDGES_L##LY(LDPC_P1_MS)                                                                \
        } else if (kind_is_faidm(KIND)) {                                                               \
            LDPC_EDGES_L##LY(LDPC_P1_FAIDM)                                                             \
        } else {                                                                                        \
            LDPC_EDGES_L##LY(LDPC_P1_FAID)                                                              \
        }                                                                                               \
        LDPC_MIN2_FINISH                                                                                \
        if (!LDPC_PRE_LATE_K && (LDPC_CV_STATIC_PRE_K ? PRE : (pre != nullptr))) {                          \
            _Pragma("unroll") for (int k = 0; k < 6; ++k) { LDPC_CV_CHECK(&pre[k * kThreads]) cv_next[k] = pre[k * kThreads]; } \
        }                                                                                               \
        uint32_t c1, c2, nthr = 0;                                                                      \
        (void)nthr;                                                                                     \
        if (KIND == KIND_NMS) {                                                                         \
            /* with the fp16 select MONO is free: for this kind it carries nms_fast */                     \
            if (LDPC_NMS_SCALE16 && LDPC_FP16_SELECT && MONO) {                                         \
                c2 = nms_scale16(min1, P.factor_1);                                                     \
                c1 = nms_scale16(min2, P.factor_2);                                                     \
            } else {                                                                                    \
                c2 = nms_scale(min1, P.factor_1);                                                       \
                c1 = nms_scale(min2, P.factor_2);                                                       \
            }                                                                                           \
        } else if (KIND == KIND_OMS) {                                                                  \
            /* the reference clips every |v| to 7 before the min search (CDecoder_OMS.cpp:374); clipping the two \
               minima afterwards is the same thing (a monotone map commutes with order statistics) */            \
            min1 = __vmins2(min1, 0x00070007u + HB);                                                    \
            min2 = __vmins2(min2, 0x00070007u + HB);                                                    \
            const uint32_t n2 = lut8(P.oms_norm[0], P.oms_norm[1], min1);                               \
            const uint32_t n1 = lut8(P.oms_norm[0], P.oms_norm[1], min2);                               \
            const uint32_t b2 = lut8(P.oms_boost[0], P.oms_boost[1], min1);                             \
            const uint32_t b1 = lut8(P.oms_boost[0], P.oms_boost[1], min2);                             \
            /* the tables hold 64 + cste (cste can be negative in OMS_MODE 0) */                        \
            c2 = __vsub2(sel32(eef, b2, n2), kP0);                                                      \
            c1 = __vsub2(sel32(eef, b1, n1), kP0);                                                      \
        } else if (kind_is_faidm(KIND)) {                                                               \
            /* min over t_j = LUT[min(|v_j|, 7)] equals LUT[min(min |v_j|, 7)] for a monotone LUT, same for the \
               second minimum; the LUT (normal or error-floor, CDecoder_FAID.cpp:712-758) is chosen per check */ \
            const uint32_t m1a = __vmins2(min1, 0x00070007u + HB), m2a = __vmins2(min2, 0x00070007u + HB); \
            uint32_t t1 = lut8(cx.lut[0][0], cx.lut[0][1], m1a), t2 = lut8(cx.lut[0][0], cx.lut[0][1], m2a); \
            uint32_t th = lut8(cx.thr[0], cx.thr[1], m1a);                                              \
            if (kind_has_ef(KIND)) {                                                                    \
                t1 = sel32(eef, lut8(cx.lut_ef[0][0], cx.lut_ef[0][1], m1a), t1);                       \
                t2 = sel32(eef, lut8(cx.lut_ef[0][0], cx.lut_ef[0][1], m2a), t2);                       \
                th = sel32(eef, lut8(cx.thr_ef[0], cx.thr_ef[1], m1a), th);                             \
            }                                                                                           \
            min1 = t1;                                                                                  \
            c2 = __vmins2(t1, 0x00070007u);                                                             \
            c1 = __vmins2(t2, 0x00070007u);                                                             \
            nthr = LDPC_FAIDM_IN_REG ? h2_sub(th + HB, HB) /* fp16(thr) */                              \
                   : HB ? th + HB /* fp16(1024 + thr), subtracted by h2_sub_sat */ : __vsub2(0u, th);   \
        } else {                                                                                        \
            if (KIND == KIND_FAID_EF) {                                                                 \
                min1 = __vmins2(min1, 0x00070007u);                                                     \
                min2 = __vmins2(min2, 0x00070007u);                                                     \
            }                                                                                           \
            c2 = __vmins2(min1, 0x00070007u);                                                           \
            c1 = __vmins2(min2, 0x00070007u);                                                           \
        }                                                                                               \
        const uint32_t P1c = __vadd2(c1, kP0 + HB), P2c = __vadd2(c2, kP0 + HB);                         \
        const uint32_t Dh = kind_fp16(KIND) ? h2_sub(P2c, P1c) : 0u;  /* fp16(cste_2 - cste_1) */         \
        (void)Dh;                                                                                       \
        const uint32_t P1big = P1c + 0x08000800u;                                                       \
        (void)P1big;                                                                                    \
        const uint32_t Qp = __vadd2(P1c + min1 * 8u, 0xF800F800u);   /* P1 + 8 min1 - 2048 */            \
        const uint32_t nmin1 = __vadd2(~min1, 0x00010001u);                                             \
        (void)Qp; (void)nmin1;                                                                          \
        /* neg_j = ~(parity ^ nonneg_j): fold the inversion into the parity word */                     \
        const uint32_t Sp = (KIND == KIND_NMS || KIND == KIND_OMS) ? (S ^ kNeg)                          \
                            : kind_is_faidm(KIND) ? ((S ^ 0x80008000u) >> 8) : (S ^ 0x80008000u);       \
        uint32_t nw = 0;                                                                                \
        if (KIND == KIND_NMS || KIND == KIND_OMS) {                                                     \
            LDPC_EDGES_L##LY(LDPC_P2_MS)                                                                \
        } else if (kind_is_faidm(KIND)) {                                                               \
            LDPC_EDGES_L##LY(LDPC_P2_FAIDM)                                                             \
        } else {                                                                                        \
            LDPC_EDGES_L##LY(LDPC_P2_FAID)                                                              \
        }                                                                                               \
        if (LDPC_PRE_LATE_K && (LDPC_CV_STATIC_PRE_K ? PRE : (pre != nullptr))) {                           \
            _Pragma("unroll") for (int k = 0; k < 6; ++k) { LDPC_CV_CHECK(&pre[k * kThreads]) cv_next[k] = pre[k * kThreads]; } \
        }                                                                                               \
    }

LDPC_FOR_EACH_LAYER(LDPC_DEF_LAYER)

// parity of the hard decisions of one row per layer (start-of-iteration syndrome)
// (Lb + hardk) has bit 15 set <=> L > 0, with hardk = 0x7FFF - bias per half (kHardK for the default bias)
#define LDPC_SYN_EDGE(j, c, s, w) X ^= __vadd2(LDPC_APP(c, LDPC_OFF(s)), hardk);
#define LDPC_SYN_LAYER(LY)                                       \
    {                                                            \
        uint32_t X = 0;                                          \
        LDPC_EDGES_L##LY(LDPC_SYN_EDGE)                          \
        chk0 |= ((X >> 15) & 1u) << LY;                          \
        chk1 |= ((X >> 31) & 1u) << LY;                          \
    }

// APP word of a frame pair: L + bias per signed 16-bit half
__host__ __device__ __forceinline__ uint32_t pack_app(int l0, int l1, int bias) {
    return ((uint32_t)(l0 + bias) & 0xFFFFu) | ((uint32_t)(l1 + bias) << 16);
}

#ifndef LDPC_HOST_EMU
}  // namespace ldpc
#include "gen_device.cuh"
namespace ldpc {

// barrier of the 256 threads that own one frame pair (named barrier 1 + slot; 0 is __syncthreads)
__device__ __forceinline__ void pair_sync(int bar_id) { asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory"); }
__device__ __forceinline__ int pair_sync_or(int bar_id, int pred) {
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %2, 0;\n\tbar.red.or.pred p, %1, 256, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(bar_id), "r"(pred)
        : "memory");
    return r;
}

// Packed hard decisions (bit n%32 of word n/32 = L[n] > 0) of both frames, optionally the 2B1C second bit.
__device__ __forceinline__ void store_hard(const uint32_t* app, uint32_t* dst0, uint32_t* dst1, int planes,
                                           int hard2_thr, int t, int bias, unsigned long long* dbg = nullptr) {
    (void)dbg;
    const int warp = t >> 5, lane = t & 31;
    const int lo_thr = bias - hard2_thr, hi_thr = bias + hard2_thr;
#pragma unroll 3
    for (int k = 0; k < kN / kThreads; ++k) {
        const uint32_t w = app[k * kThreads + t];
        const int l0 = (int)(int16_t)(w & 0xFFFFu), l1 = (int)(int16_t)(w >> 16);  // halves are signed (bias_of)
        const uint32_t h0 = __ballot_sync(0xFFFFFFFFu, l0 > bias);
        const uint32_t h1 = __ballot_sync(0xFFFFFFFFu, l1 > bias);
        if (lane == 0) {
            LDPC_CHECK(dbg, k * 8 + warp < kHW, DBG_HARD, k * 8 + warp);
            dst0[k * 8 + warp] = h0;
            dst1[k * 8 + warp] = h1;
        }
        if (planes > 1) {
            const uint32_t g0 = __ballot_sync(0xFFFFFFFFu, l0 >= hi_thr || l0 <= lo_thr);
            const uint32_t g1 = __ballot_sync(0xFFFFFFFFu, l1 >= hi_thr || l1 <= lo_thr);
            if (lane == 0) {
                dst0[kHW + k * 8 + warp] = g0;
                dst1[kHW + k * 8 + warp] = g1;
            }
        }
    }
}

template <int KIND, bool MONO>
__global__ void __launch_bounds__(kThreads * kPairsPerCta, 2 / kPairsPerCta) decode_pair_kernel(const DecParams P) {
    extern __shared__ uint32_t smem_all[];
    // which frame pair of this CTA; taken from a warp vote so that ptxas knows it is warp-uniform and keeps the pair's
    // shared-memory base in a uniform register (LDS [R + UR + imm]) instead of spending a vector register and an add
    const int slot = kPairsPerCta > 1 ? (int)__any_sync(0xFFFFFFFFu, threadIdx.x >= kThreads) : 0;
    const int bar = 1 + slot;
    // [kN] one word per code bit: frame 2p in the low half, 2p+1 in the high half; then the pair's message words
    uint32_t* const app_pair = smem_all + (size_t)slot * pair_smem_words(KIND);
    uint32_t* const app = smem_all;  // base for LDPC_APP: the offsets of LDPC_OFF already contain the pair's base (pbase)
    constexpr int kCvSmemLayers = cv_smem_layers(KIND);
    constexpr int kB = bias_of(KIND);
    constexpr uint32_t hardk = hardk_of(KIND);
    (void)hardk;
    const int t = threadIdx.x % kThreads;

    __shared__ int s_err_all[kPairsPerCta][2][2];  // per-frame unsatisfied-row counts, double-buffered by iteration parity
    int(&s_err)[2][2] = s_err_all[slot];

    const uint32_t pbase = (uint32_t)slot * (uint32_t)(pair_smem_words(KIND) * sizeof(uint32_t));  // multiple of 1024
    const uint32_t rr = (uint32_t)t * 4u + pbase;  // byte offset of row t inside block column 0 of this pair's APP array
    // [kCvSmemLayers][6][kThreads] message words of the "cold" layers, right behind the APP array: same register as rr
    uint32_t* const cvs = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(smem_all) + rr) + kN;
    // Work loop: one pass with LDPC_PERSISTENT 0 (the CTA's own two frame pairs), ticket-driven otherwise (see the switch).
#if LDPC_PERSISTENT
    __shared__ int s_ticket[kPairsPerCta];
    __shared__ volatile int s_go;
    if (threadIdx.x == 0) s_go = 0;
    __syncthreads();
#endif
    const int n_pairs = P.n_frames >> 1;
    bool first_item = true;
    for (;;) {
    // all threads of the half are past their last access to the previous item's APP words when they arrive here
#if LDPC_PERSISTENT
    if (t == 0) s_ticket[slot] = (int)atomicAdd(P.work_counter, 1u);
    pair_sync(bar);
    const int pair = s_ticket[slot];
#else
    const int pair = first_item ? (int)blockIdx.x * kPairsPerCta + slot : n_pairs;
#endif
    if (pair >= n_pairs) {
#if LDPC_PERSISTENT
        if (slot == 0 && t == 0) s_go = 1;  // the other half must never wait for a half that has no work
#endif
        break;
    }
    const int f0 = pair * 2;
    const int group = f0 >> 5;

    // ---- load channel LLRs (CLDPC.cpp:234-272) ----
    if (P.gen_enable) {
        // Fused producer: the pair's 256 threads synthesise the two frames -- map, Philox AWGN, demap,
        // de-interleave, 4-bit quantise (CSimulate.cpp:111-132) -- straight into the APP array; no LLR ever touches HBM.
        // Bit-identical to generate_kernel + the load below (same Philox stream, same float operations).  Its FMA- and
        // XU-pipe work overlaps the ALU-bound decoding of the other pair on the SM.
        const GenCore& G = P.gen;
        const int pairs_per_frame = kN / G.mod / 2;
        const int fg = f0 & 31;
        const uint64_t gf = G.first_frame + (uint64_t)f0;
        const int txg = tx_group_of(G, group);
        if (G.fast_i1) {
            // identity interleaver: a thread makes the same symbol pair of BOTH frames and stores whole APP words
#define LDPC_GEN_FAST(MOD)                                                                         \
    for (int sp = t; sp < pairs_per_frame; sp += kThreads) {                                       \
        float re[2], im[2];                                                                        \
        int q0[2 * MOD], q1[2 * MOD];                                                              \
        gen_symbol_pair_i1<MOD>(G, txg, fg, gf, sp, re, im);                                     \
        demap_quant_pair<MOD>(re, im, G.scale, G.qbits, q0);                                                \
        gen_symbol_pair_i1<MOD>(G, txg, fg + 1, gf + 1, sp, re, im);                             \
        demap_quant_pair<MOD>(re, im, G.scale, G.qbits, q1);                                                \
        uint4* dst = reinterpret_cast<uint4*>(app_pair + 2 * MOD * sp);                            \
        _Pragma("unroll") for (int i = 0; i < MOD / 2; ++i)                                        \
            dst[i] = make_uint4(pack_app(q0[4 * i], q1[4 * i], kB), pack_app(q0[4 * i + 1], q1[4 * i + 1], kB), \
                                pack_app(q0[4 * i + 2], q1[4 * i + 2], kB), pack_app(q0[4 * i + 3], q1[4 * i + 3], kB)); \
    }
            if (G.mod == 2) { LDPC_GEN_FAST(2) } else if (G.mod == 4) { LDPC_GEN_FAST(4) } else if (G.mod == 6) { LDPC_GEN_FAST(6) } else { LDPC_GEN_FAST(8) }
#undef LDPC_GEN_FAST
        } else {
            uint16_t* const app16 = reinterpret_cast<uint16_t*>(app_pair);
#pragma unroll 1
            for (int fsel = 0; fsel < 2; ++fsel) {
#pragma unroll 1
                for (int sp = t; sp < pairs_per_frame; sp += kThreads) {
                    float re[2], im[2];
                    gen_symbol_pair(G, txg, fg + fsel, gf + fsel, sp, re, im);
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float llr[kMaxMod];
                        demap_symbol(re[k], im[k], G.mod, llr);
                        for (int b = 0; b < G.mod; ++b) {
                            const int src = (2 * sp + k) * G.mod + b;
                            const int i = src / G.I, j = src - i * G.I;
                            const int dst = j * (kN / G.I) + i;  // code-bit index (CModulate.cpp:161-172)
                            app16[2 * dst + fsel] = (uint16_t)(quant_cfg(llr[b], G.scale, G.qbits) + kB);
                        }
                    }
                }
            }
        }
    } else if (P.exp_noload) {
        // experiment (LDPC_B200_EXP_NOLOAD): no LLR load at all, to time what the load phase costs; results are meaningless
        for (int n = t; n < kN; n += kThreads) app_pair[n] = pack_app((n * 7 + t) % 13 - 6, (n * 5) % 11 - 5, kB);
    } else if (P.llr) {
        const int fg = f0 & 31;
        const int8_t* base = P.llr + (size_t)group * 32 * kN;
        const uint32_t* i0 = reinterpret_cast<const uint32_t*>(base + (size_t)fg * kK);
        const uint32_t* i1 = reinterpret_cast<const uint32_t*>(base + (size_t)(fg + 1) * kK);
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(base + (size_t)32 * kK + (size_t)fg * kM);
        const uint32_t* p1 = reinterpret_cast<const uint32_t*>(base + (size_t)32 * kK + (size_t)(fg + 1) * kM);
        // 4416 words per frame: every load of the thread (18 per frame) is in flight at once -- one exposure to the HBM latency
        // per CTA instead of one per batch -- and the expansion is 4 instructions per code bit pair: the bytes are made
        // unsigned (L + 128, one XOR per word), PRMT spreads a byte of each frame into the halves of a word, and one
        // VIADDMNMX + one VIMNMX add the bias and clamp.  (NOLOAD experiment, profiles/r02_nms_ab_exp4.log: the load phase
        // was 6.6 % of the kernel, 2.6 % of it instructions.)
        // Full int8 range, exactly as the reference's 8-bit saturating arithmetic treats it: a code bit's first V2C is
        // v = sat8(L - 0) (all messages start at 0) and every later L is within [-31, 31].  Below, v is clamped at -31 by
        // every decoder (CLDPC.cpp:330); above, the min-sum decoders leave v unclamped, but the minima are capped at 31 and
        // L' = min(v + c, 31) with c >= -7, so every L >= 39 acts like 39; the FAID decoders clamp v at +31 as well.
        // (tests/test_gpu_decode.py::test_full_int8_range, against the reference)
        constexpr int kWords = kN / 4, kPer = (kWords + kThreads - 1) / kThreads;  // 18
        constexpr int kHiL = (KIND == KIND_NMS || KIND == KIND_OMS) ? 39 : 31;
        constexpr uint32_t cadd = ((uint32_t)(kB - 128) & 0xFFFFu) * 0x00010001u;
        constexpr uint32_t chi = ((uint32_t)(kB + kHiL) & 0xFFFFu) * 0x00010001u, clo = ((uint32_t)(kB - 31) & 0xFFFFu) * 0x00010001u;
        uint32_t a[kPer], b[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int q = t + k * kThreads;
            a[k] = b[k] = 0;
            LDPC_CHECK(P.dbg, fg + 1 < 32 && (size_t)group * 32 + fg + 1 < (size_t)P.n_frames, DBG_LLR, f0);
            if (q < kK / 4) { a[k] = __ldg(i0 + q); b[k] = __ldg(i1 + q); }
            else if (q < kWords) { a[k] = __ldg(p0 + q - kK / 4); b[k] = __ldg(p1 + q - kK / 4); }
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int q = t + k * kThreads;
            if (q < kWords) {
                const uint32_t ax = a[k] ^ 0x80808080u, bx = b[k] ^ 0x80808080u;  // bytes = L + 128, unsigned
                const uint32_t lo = __byte_perm(ax, bx, 0x5140), hi = __byte_perm(ax, bx, 0x7362);  // [a0 b0 a1 b1], [a2 b2 a3 b3]
                uint4 o;
                o.x = __vmaxs2(__viaddmin_s16x2(__byte_perm(lo, 0u, 0x4140), cadd, chi), clo);
                o.y = __vmaxs2(__viaddmin_s16x2(__byte_perm(lo, 0u, 0x4342), cadd, chi), clo);
                o.z = __vmaxs2(__viaddmin_s16x2(__byte_perm(hi, 0u, 0x4140), cadd, chi), clo);
                o.w = __vmaxs2(__viaddmin_s16x2(__byte_perm(hi, 0u, 0x4342), cadd, chi), clo);
                reinterpret_cast<uint4*>(app_pair)[q] = o;
            }
        }
    } else {
        // native layout: frame-major, two 4-bit two's-complement LLRs per byte (low nibble = even code bit)
        const uint32_t* n0 = reinterpret_cast<const uint32_t*>(P.llr_packed + (size_t)f0 * (kN / 2));
        const uint32_t* n1 = reinterpret_cast<const uint32_t*>(P.llr_packed + (size_t)(f0 + 1) * (kN / 2));
        // 2208 words per frame, 9 per thread and frame, all in flight at once; expansion as in the int8 loader: nibbles made
        // unsigned (L + 8, one XOR), even / odd nibbles separated into bytes, PRMT pairs the two frames' bytes into halves
        constexpr int kWords = kN / 8, kPer = (kWords + kThreads - 1) / kThreads;  // 9
        constexpr uint32_t cadd = ((uint32_t)(kB - 8) & 0xFFFFu) * 0x00010001u;
        uint32_t a[kPer], b[kPer];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int q = t + k * kThreads;
            a[k] = b[k] = 0;
            LDPC_CHECK(P.dbg, f0 + 1 < P.n_frames, DBG_LLR, f0);
            if (q < kWords) { a[k] = __ldg(n0 + q); b[k] = __ldg(n1 + q); }
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int q = t + k * kThreads;
            if (q < kWords) {
                const uint32_t ax = a[k] ^ 0x88888888u, bx = b[k] ^ 0x88888888u;      // nibbles = L + 8, unsigned
                const uint32_t ae = ax & 0x0F0F0F0Fu, ao = (ax >> 4) & 0x0F0F0F0Fu;   // code bits 0,2,4,6 / 1,3,5,7 of the word
                const uint32_t be = bx & 0x0F0F0F0Fu, bo = (bx >> 4) & 0x0F0F0F0Fu;
                const uint32_t e01 = __byte_perm(ae, be, 0x5140), e23 = __byte_perm(ae, be, 0x7362);  // [a0 b0 a2 b2], [a4 b4 a6 b6]
                const uint32_t o01 = __byte_perm(ao, bo, 0x5140), o23 = __byte_perm(ao, bo, 0x7362);  // [a1 b1 a3 b3], [a5 b5 a7 b7]
                uint4 u0, u1;
                u0.x = __vadd2(__byte_perm(e01, 0u, 0x4140), cadd);  // bit 0
                u0.y = __vadd2(__byte_perm(o01, 0u, 0x4140), cadd);  // bit 1
                u0.z = __vadd2(__byte_perm(e01, 0u, 0x4342), cadd);  // bit 2
                u0.w = __vadd2(__byte_perm(o01, 0u, 0x4342), cadd);  // bit 3
                u1.x = __vadd2(__byte_perm(e23, 0u, 0x4140), cadd);  // bit 4
                u1.y = __vadd2(__byte_perm(o23, 0u, 0x4140), cadd);  // bit 5
                u1.z = __vadd2(__byte_perm(e23, 0u, 0x4342), cadd);  // bit 6
                u1.w = __vadd2(__byte_perm(o23, 0u, 0x4342), cadd);  // bit 7
                reinterpret_cast<uint4*>(app_pair)[2 * q] = u0;
                reinterpret_cast<uint4*>(app_pair)[2 * q + 1] = u1;
            }
        }
    }
    pair_sync(bar);
#if LDPC_DEBUG_BOUNDS
    // negative control of the checker: the offset one word past this pair's APP array is checked (and not dereferenced)
    if (P.exp_fault && t == 0 && pair == 0) (void)ldpc_app_checked(app, pbase + kN * 4u, pbase, P.dbg);
#endif
    for (int n = kN - P.puncture_tail + t; n < kN; n += kThreads) app_pair[n] = pack_app(0, 0, kB);

    // messages start at 0: stored nibble = m + 8
    uint32_t cvr[LDPC_MB - kCvSmemLayers][6];  // register-resident layers (kCvSmemLayers < LDPC_MB)
    uint32_t cva[6], cvb[6];                   // ping-pong staging of the shared-memory-resident layers
#pragma unroll
    for (int l = 0; l < LDPC_MB - kCvSmemLayers; ++l)
#pragma unroll
        for (int k = 0; k < 6; ++k) cvr[l][k] = 0x88888888u;
#pragma unroll
    for (int k = 0; k < kCvSmemLayers * 6; ++k) cvs[k * kThreads] = 0x88888888u;
#pragma unroll
    for (int k = 0; k < 6; ++k) cva[k] = cvb[k] = 0x88888888u;
    if (t < 4) (&s_err[0][0])[t] = 0;
    pair_sync(bar);

#if LDPC_PERSISTENT
    if (kPairsPerCta > 1 && first_item && slot == 1 && P.max_iter >= 2 && !P.no_skew) {
        // first item of the second half: start once the first half is half-way through its own first item
        if (t == 0)
            while (!s_go) __nanosleep(256);
        pair_sync(bar);
    }
#endif
    int fz0 = 0, fz1 = 0;  // 1 + first iteration index with zero syndrome
    IterCtx cx;
    cx.chk0 = cx.chk1 = 0;
    cx.lane_ok = 0;
    cx.special_active = 0;
    bool stopped = false;

    for (int it = 1; it <= P.max_iter; ++it) {
        const int remaining = P.max_iter - it;
#if LDPC_PERSISTENT
        if (kPairsPerCta > 1 && first_item && slot == 0 && t == 0 && it == (P.max_iter >> 1) + 1) s_go = 1;
#endif
        if (KIND != KIND_NMS) {
            // ---- start-of-iteration syndrome (CDecoder_OMS.cpp:102-136, CDecoder_FAID.cpp:294-343) ----
            int(&se)[2] = s_err[it & 1];
            uint32_t chk0 = 0, chk1 = 0;
            LDPC_FOR_EACH_LAYER(LDPC_SYN_LAYER)
            const int e0 = __reduce_add_sync(0xFFFFFFFFu, __popc(chk0));
            const int e1 = __reduce_add_sync(0xFFFFFFFFu, __popc(chk1));
            if ((t & 31) == 0) {
                if (e0) atomicAdd(&se[0], e0);
                if (e1) atomicAdd(&se[1], e1);
            }
            if (KIND == KIND_FAID_ER) app_pair[unsat_word_offset(KIND) + t] = chk0 | (chk1 << 16);
            pair_sync(bar);
            const int err0 = min(se[0], P.err_sat), err1 = min(se[1], P.err_sat);
            if (t < 2) s_err[(it + 1) & 1][t] = 0;  // next iteration's buffer; its atomics come >= 12 barriers later
            const int z0 = err0 == 0, z1 = err1 == 0;
            if (z0 && !fz0) fz0 = it;
            if (z1 && !fz1) fz1 = it;
            if (z0 | z1) {
                // snapshot of the hard decisions: the group's stop iteration may turn out to be this one
                LDPC_CHECK(P.dbg, f0 + 1 < P.n_frames && it - 1 < P.max_iter && P.snap != nullptr, DBG_SNAP, f0);
                uint32_t* s0 = P.snap + (((size_t)f0 * P.max_iter + (it - 1)) * P.planes) * kHW;
                uint32_t* s1 = P.snap + (((size_t)(f0 + 1) * P.max_iter + (it - 1)) * P.planes) * kHW;
                // Thread 0 publishes this pair's frames first, so that the atomic's round trip overlaps the snapshot.
                // Opportunistic stop: has the whole group been seen converged at this or an earlier iteration?  The
                // older counters are read by `it - 1` different threads at once (one L2 round trip, not `it`).
                unsigned int* cnt = P.grp_cnt + (size_t)group * P.max_iter;
                int seen = 0;
                if (t == 0) seen = atomicAdd(&cnt[it - 1], (unsigned)(z0 + z1)) + (unsigned)(z0 + z1) == 32u;
                else if (t < it) seen = *reinterpret_cast<volatile unsigned int*>(cnt + (t - 1)) == 32u;
                store_hard(app_pair, s0, s1, P.planes, P.hard2_thr, t, kB, P.dbg);
                if (pair_sync_or(bar, seen)) { stopped = true; break; }
            }
            cx.chk0 = chk0;
            cx.chk1 = chk1;
            if (KIND == KIND_OMS) {
                cx.lane_ok = expand2((unsigned)err0 < (unsigned)P.oms_floor_err, (unsigned)err1 < (unsigned)P.oms_floor_err);
                cx.special_active = remaining <= P.oms_floor_iter;
            } else {
                cx.lane_ok = expand2(err0 < P.ef_floor_err, err1 < P.ef_floor_err);
                cx.special_active = kind_has_ef(KIND) && remaining <= P.ef_floor_iter;
            }
        }
        if (kind_is_faid(KIND)) {
            const int li = (it < 6 ? it : 6) - 1;  // switch (nb_iteration - nombre_iterations), CDecoder_FAID.cpp:760-781
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                cx.thr[k] = P.luts.thr[li][k];
                cx.thr_ef[k] = P.luts.thr_ef[li][k];
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                cx.lut[w][0] = P.luts.lut[li][w][0];
                cx.lut[w][1] = P.luts.lut[li][w][1];
                cx.lut_ef[w][0] = P.luts.lut_ef[li][w][0];
                cx.lut_ef[w][1] = P.luts.lut_ef[li][w][1];
            }
        }
// layer LY < kCvSmemLayers works on the staging buffer (LY & 1) that the previous layer prefetched
#define LDPC_CV_CUR(LY) ((LY) < kCvSmemLayers ? (((LY) & 1) ? cvb : cva) : cvr[(LY) < kCvSmemLayers ? 0 : (LY) - kCvSmemLayers])
#define LDPC_NEXT(LY) (((LY) + 1) % LDPC_MB)
#define LDPC_RUN_LAYER(LY)                                                                              \
    layer_##LY<KIND, MONO, ((LY) < kCvSmemLayers), (LDPC_NEXT(LY) < kCvSmemLayers)>(app, rr, pbase, LDPC_CV_CUR(LY), \
                           (LY) < kCvSmemLayers ? cvs + (LY) * 6 * kThreads : nullptr,                  \
                           LDPC_NEXT(LY) < kCvSmemLayers ? cvs + LDPC_NEXT(LY) * 6 * kThreads : nullptr, \
                           (LDPC_NEXT(LY) & 1) ? cvb : cva, cx, P);                                     \
    pair_sync(bar);
        LDPC_FOR_EACH_LAYER(LDPC_RUN_LAYER)
#undef LDPC_RUN_LAYER
#undef LDPC_CV_CUR
#undef LDPC_NEXT
    }

    if (KIND == KIND_NMS && P.direct_bytes) {
        // hard decision + "inverse transpose" (CLDPC.cpp:2268-2270, CTool.cpp:291-575) straight from the APP array.
        // A warp takes 128 code bits per step: lane l reads the uint4 of bits 4l..4l+3 (consecutive 16-byte words, no bank
        // conflict -- one lane per 16 CONSECUTIVE bits was a 64-byte stride, 4-way conflicts) and stores 4 bytes per frame.
        uint32_t* o0 = reinterpret_cast<uint32_t*>(P.direct_bytes + (size_t)f0 * kN);
        uint32_t* o1 = reinterpret_cast<uint32_t*>(P.direct_bytes + (size_t)(f0 + 1) * kN);
        for (int u = t >> 5; u < kN / 128; u += kThreads / 32) {
            const int idx = 32 * u + (t & 31);
            LDPC_CHECK(P.dbg, idx < kN / 4 && f0 + 1 < P.n_frames, DBG_HARD, idx);
            const uint4 w = reinterpret_cast<const uint4*>(app_pair)[idx];
            // (Lb + hardk) has bit 15 of its half set <=> L > 0: the decisions of four code bits sit in bit 7 of bytes 1 (frame 0)
            // and 3 (frame 1) of four words; two PRMT levels gather them into one word per frame
            const uint32_t tx = __vadd2(w.x, hardk), ty = __vadd2(w.y, hardk), tz = __vadd2(w.z, hardk), tw = __vadd2(w.w, hardk);
            const uint32_t pxy = __byte_perm(tx, ty, 0x7351), pzw = __byte_perm(tz, tw, 0x7351);  // [x.b1 y.b1 x.b3 y.b3], [z.b1 w.b1 z.b3 w.b3]
            const uint32_t b0 = (__byte_perm(pxy, pzw, 0x5410) >> 7) & 0x01010101u;
            const uint32_t b1 = (__byte_perm(pxy, pzw, 0x7632) >> 7) & 0x01010101u;
            o0[idx] = b0;
            o1[idx] = b1;
        }
    } else if (KIND == KIND_NMS && P.direct_packed) {
        store_hard(app_pair, P.direct_packed + (size_t)f0 * kHW, P.direct_packed + (size_t)(f0 + 1) * kHW, 1, P.hard2_thr, t, kB, P.dbg);
    } else if (!stopped) {
        LDPC_CHECK(P.dbg, f0 + 1 < P.n_frames && P.final_hard != nullptr, DBG_HARD, f0);
        uint32_t* d0 = P.final_hard + (size_t)f0 * P.planes * kHW;
        uint32_t* d1 = P.final_hard + (size_t)(f0 + 1) * P.planes * kHW;
        store_hard(app_pair, d0, d1, P.planes, P.hard2_thr, t, kB, P.dbg);
    }
    if (t == 0 && P.first_zero) {
        P.first_zero[f0] = fz0;
        P.first_zero[f0 + 1] = fz1;
    }
#if LDPC_PERSISTENT
    if (slot == 0 && t == 0) s_go = 1;  // covers an early group stop before the half-way iteration
#endif
    first_item = false;
    }  // next frame pair
}

#endif  // !LDPC_HOST_EMU

}  // namespace ldpc
