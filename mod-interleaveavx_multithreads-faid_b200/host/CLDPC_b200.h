// CLDPC_b200.h -- CLDPC-shaped C++ shim over the C-ABI (include/ldpc_b200.h).
//
// Mirrors the reference's operator interface for the hot path (CLDPC.h:110-171): same public buffer members, same
// argument-less Decode*() entry points, same int BFiter returns, same Statistic result, so that a CSimulate::Run-style
// loop (CSimulate.cpp:103-169) reads exactly like the reference's.  Differences, all deliberate:
//   * Factor_1/Factor_2/scale come from the Parameter_Simulation given to Initial() (the reference re-reads
//     ./Profile.txt inside every call, CLDPC.cpp:216-217);
//   * errors throw std::runtime_error instead of exit(EXIT_FAILURE) + getchar();
//   * n_groups >= 1 groups of 32 frames can be decoded per call (the buffers are n_groups times larger).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "ldpc_b200.h"

// When this header is included AFTER the reference's own headers (a maintainer swapping `CLDPC` for `CLDPC_B200` inside the
// reference tree; oracle/ref_build/dropin_csimulate.cpp does exactly that), the reference's Parameter_Simulation
// (CTool.h:23-39), ReadProfile (CTool.h:47) and Statistic (CLDPC.h:103-108) are used as they are.
#ifndef CTOOL_H
struct Parameter_Simulation {  // CTool.h:23-39, same field names
    float snr_start, snr_pass, snr_end, scale;
    int decode_method, Max_Iteration, mod_type, interleavemod_type, Factor_1, Factor_2, nb_frames, Z;
};
#endif
// CTool.cpp:588-621 with an explicit path (the reference's ReadProfile(p) reads ./Profile.txt)
void ReadProfile(Parameter_Simulation* p, const char* path);
#ifndef CTOOL_H
inline void ReadProfile(Parameter_Simulation* p) { ReadProfile(p, "Profile.txt"); }
#endif

#ifndef CLDPC_H
struct Statistic {  // CLDPC.h:103-108
    unsigned long ErrorFrame, ErrorBits, LT3ErrBitFrame;
};
#endif

class CLDPC_B200 {
public:
    double m_Rate = 0.8444444;  // CLDPC.cpp:4780
    int8_t* inputBits = nullptr;    // [n_groups][32*K]
    int8_t* outputBits = nullptr;   // [n_groups][32*N] two-region layout
    int8_t* decodedBits = nullptr;  // [n_groups][32*N] frame-major 0/1
    int8_t* fixInput = nullptr;     // [n_groups][32*N] two-region layout
    int nb_iteration = 0, m_M = LDPC_B200_M, m_N = LDPC_B200_N, m_K = LDPC_B200_K, m_frame = 32;
    std::vector<int32_t> its_per_group, bf_iters, conv_iter;  // extra observables of the last Decode*()

    CLDPC_B200() = default;
    ~CLDPC_B200();
    CLDPC_B200(const CLDPC_B200&) = delete;
    CLDPC_B200& operator=(const CLDPC_B200&) = delete;

    // CLDPC::Initial(nb_frame, MaxIteration) + the Profile fields the decoders use; lut_variant: LDPC_B200_LUT_*
    void Initial(const Parameter_Simulation& p, int n_groups = 1, int device = 0, int lut_variant = -1);
    // The reference's own signature (CLDPC.cpp:4772): Factor_1 / Factor_2 / scale / modType come from ./Profile.txt, which
    // the reference re-reads inside every Decode*() (CLDPC.cpp:216-217) and this shim reads once, here.
    void Initial(int nb_frame, int MaxIteration);

    void GenMsgSeq();                      // CLDPC.cpp:60-66 (rand()%2)
    void Encode();                         // CLDPC.cpp:68-126 (systematic encoder derived from H)
    void FakeEncoder(const int* codeword); // CLDPC.cpp:163-207 (same codeword in all lanes)
    void FakeEncoder();                    // ... with the shipped CodeWord_sym, which is all-zero (Codeword.h:4)

    void Decode();            // NMS            CLDPC.cpp:214
    void Decode_OMS();        //                CDecoder_OMS.cpp:13
    void Decode_FAID();       // FAID + DTBF    CDecoder_FAID.cpp:176
    int Decode_OMSBF();       // returns BFiter CDecoder_OMSBF.cpp:13
    int Decode_OMS_DTBF();    // returns BFiter CDecoder_OMS_DTBF.cpp:18
    void Decode_FAID_2B1C();  //                CDecoder_FAID_2B1C.cpp:96
    int DecodeDispatch(int decode_method);  // the switch of CSimulate.cpp:136-164

    void float2LimitChar_4bit(int8_t* output, const float* input, float scale, int length);  // CLDPC.cpp:4524
    // fused producer (CSimulate.cpp:111-132): outputBits -> fixInput at Eb/N0, Philox noise
    void GenerateNoisyBlock(float Eb_N0, uint64_t seed, uint64_t first_frame_index);
    Statistic CalculateErrors();  // CLDPC.cpp:4819 (info bits only)
    // the reference's signature; its arguments only feed the error-frame dumps (CLDPC.cpp:4877-4991)
    Statistic CalculateErrors(float* bpskinput, int8_t* charinput, int collectflag);

    ldpc_b200_handle* handle() { return h_; }

private:
    void ensure(int method);
    int run(int method);
    ldpc_b200_handle* h_ = nullptr;
    int h_method_ = -1;
    int n_groups_ = 0, device_ = 0, lut_ = -1;
    Parameter_Simulation p_{};
};
