/* ldpc_oracle.h -- CPU restatement of the reference's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this; the product (libldpc_b200.so) never does and
 * has no CPU fallback.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_vs_reference.py against the
 * reference's own translation units compiled unmodified into oracle/_ref/ (see oracle/Makefile) on
 * dumped inputs, and against the committed fixtures under tests/golden/ (which were produced by that
 * reference build with tools/make_golden.py).  The only in-tree known-answer vector of the reference,
 * the "50G PON NS NP" codeword (Codeword.h:6-460), pins H and the encoder (tools/gen_code_tables.py).
 * NOT pinned: the MKL MT2203 Gaussian stream of the BPSK channel (CChannel.cpp:49,102-109) -- MKL is not
 * vendored and BPSK is outside the BASELINE configs.
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H

#include <stdint.h>

#include "ldpc_b200.h" /* for ldpc_b200_config (the by-value configuration struct) only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ldpc_oracle_info {
    int32_t iters_executed; /* min-sum iterations whose check-node pass ran */
    int32_t bf_iters;       /* BF iterations executed (BFiter), 0 when no BF stage */
    int32_t conv_iter[32];  /* per lane: completed iterations at the first start-of-iteration zero syndrome, else -1 */
    int32_t errsum_n;       /* number of logged iterations (= iters_executed for methods 1..5) */
    uint8_t errsum_log[64][32]; /* per executed iteration: per-lane error_sum after the early-stop test */
} ldpc_oracle_info;

/* The reference's shipped constants (restated independently of the product's ldpc_b200_default_config). */
void ldpc_oracle_default_config(ldpc_b200_config* cfg, int decode_method, int lut_variant);

/* One group of 32 frames: fixInput int8[32*N] (reference layout) -> decodedBits int8[32*N] (0/1, frame-major). */
int ldpc_oracle_decode(const ldpc_b200_config* cfg, const int8_t* fixInput, int8_t* decodedBits, ldpc_oracle_info* info);

/* CLDPC::float2LimitChar_4bit (CLDPC.cpp:4524-4582) */
void ldpc_oracle_quantize_4bit(int8_t* out, const float* in, float scale, int64_t length);
int ldpc_oracle_quantize_bits(int8_t* out, const float* in, float scale, int64_t length, int bits);

/* CTool.cpp:9-289 / 293-575: n x 32 byte transposes; the inverse applies (x > 0). */
void ldpc_oracle_transpose(const int8_t* src, int8_t* dst, int n);
void ldpc_oracle_itranspose(const int8_t* src, int8_t* dst, int n);

/* CModulate (CModulate.cpp:95-362), one group of 32 frames.
 * outputBits int8[32*N] two-region layout -> symbols complex64[32*N/mod_type] (interleaved re,im). */
int ldpc_oracle_modulate(const int8_t* outputBits, int mod_type, int interleave, float* symbols);
/* symbols -> DemodSeq float[32*N] (optional) -> DeInterLeaveSeq float[32*N] in the two-region layout. */
int ldpc_oracle_demodulate(const float* symbols, int mod_type, int interleave, float* demod, float* deint);

/* CModulate::BPSKModulation (CModulate.cpp:363-370): outputBits int8[32*N] -> float[32*N] = 2 b - 1, same (two-region) order.
 * The BPSK receive side is the quantiser applied to the noisy amplitudes (CSimulate.cpp:121-124). */
void ldpc_oracle_bpsk_modulate(const int8_t* outputBits, float* symbols);

/* CChannel::AWGNChannel with the 3-LCG uniform + Box-Muller (CChannel.cpp:71-97).  state[3] = IX,IY,IZ. */
void ldpc_oracle_awgn(const float* in_symbols, float* out_symbols, int64_t n_symbols, float sigma, uint64_t state[3]);
/* sigma of CSimulate::Configure (CSimulate.cpp:67-75) */
float ldpc_oracle_sigma(float ebn0_db, int mod_type, double rate);

/* Systematic encoder pinned by H + the golden codeword: info int8[K] (one frame) -> codeword int8[N]. */
void ldpc_oracle_encode_frame(const int8_t* info, int8_t* codeword);
/* Syndrome weight of one frame (0 = codeword). */
int ldpc_oracle_syndrome_weight(const int8_t* codeword);
/* Group-level encode in the reference layouts: inputBits int8[32*K] -> outputBits int8[32*N] (two regions). */
void ldpc_oracle_encode_group(const int8_t* inputBits, int8_t* outputBits);

/* CLDPC::CalculateErrors (CLDPC.cpp:4819-4995): stats3 = {ErrorFrame, ErrorBits, LT3ErrBitFrame} for one group. */
void ldpc_oracle_calc_errors(const int8_t* inputBits, const int8_t* decodedBits, uint64_t stats3[3]);

/* CPU baseline for bench.py ("port" kind): n_threads threads each decoding the given groups round-robin for
 * at least min_seconds; returns frames per second. */
double ldpc_oracle_bench_decode(const ldpc_b200_config* cfg, int n_threads, double min_seconds, const int8_t* groups,
                                int n_groups, int64_t* frames_done);

#ifdef __cplusplus
}
#endif
#endif
