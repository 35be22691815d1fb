#include "CLDPC_b200.h"

#include <cstdlib>
#include <cstring>
#include <fstream>

static void check(int rc, const char* what) {
    if (rc != LDPC_B200_OK) throw std::runtime_error(std::string(what) + ": " + ldpc_b200_last_error());
}

void ReadProfile(Parameter_Simulation* p, const char* path) {
    ldpc_b200_config c;
    std::memset(&c, 0, sizeof c);
    check(ldpc_b200_read_profile(path, &c, -1), "ReadProfile");
    p->snr_start = c.snr_start; p->snr_pass = c.snr_pass; p->snr_end = c.snr_end; p->scale = c.scale;
    p->decode_method = c.decode_method; p->Max_Iteration = c.max_iteration; p->mod_type = c.mod_type;
    p->interleavemod_type = c.interleave_mod_type; p->Factor_1 = c.factor_1; p->Factor_2 = c.factor_2;
    p->nb_frames = c.nb_frames; p->Z = c.Z;
}

CLDPC_B200::~CLDPC_B200() {
    if (h_) ldpc_b200_destroy(h_);
    ldpc_b200_host_free(inputBits);
    ldpc_b200_host_free(outputBits);
    ldpc_b200_host_free(decodedBits);
    ldpc_b200_host_free(fixInput);
}

void CLDPC_B200::Initial(const Parameter_Simulation& p, int n_groups, int device, int lut_variant) {
    p_ = p;
    n_groups_ = n_groups;
    device_ = device;
    lut_ = lut_variant;
    nb_iteration = p.Max_Iteration;
    const size_t g = (size_t)n_groups;
    // pinned so that the staged copies of Decode*() run at link speed
    check(ldpc_b200_host_alloc((void**)&inputBits, g * 32 * LDPC_B200_K), "alloc inputBits");
    check(ldpc_b200_host_alloc((void**)&outputBits, g * 32 * LDPC_B200_N), "alloc outputBits");
    check(ldpc_b200_host_alloc((void**)&decodedBits, g * 32 * LDPC_B200_N), "alloc decodedBits");
    check(ldpc_b200_host_alloc((void**)&fixInput, g * 32 * LDPC_B200_N), "alloc fixInput");
    its_per_group.assign(g, 0);
    bf_iters.assign(g, 0);
    conv_iter.assign(g * 32, -1);
    ensure(p.decode_method);
}

void CLDPC_B200::Initial(int nb_frame, int MaxIteration) {
    if (nb_frame != 32) throw std::runtime_error("CLDPC_B200::Initial: nb_frame must be 32");
    Parameter_Simulation p;
    std::memset(&p, 0, sizeof p);
    ReadProfile(&p, "Profile.txt");
    p.Max_Iteration = MaxIteration;
    Initial(p, 1, 0, -1);
}

// one engine handle per DecodeMethod in use (the kernels and constants differ per method)
void CLDPC_B200::ensure(int method) {
    if (h_ && h_method_ == method) return;
    if (h_) ldpc_b200_destroy(h_);
    h_ = nullptr;
    ldpc_b200_config c;
    check(ldpc_b200_default_config(&c, method, lut_), "default_config");
    c.max_iteration = nb_iteration;
    c.mod_type = p_.mod_type;
    c.interleave_mod_type = p_.interleavemod_type;
    c.factor_1 = p_.Factor_1;
    c.factor_2 = p_.Factor_2;
    c.scale = p_.scale;
    c.device = device_;
    c.chunk_groups = n_groups_ < 256 ? n_groups_ : 256;
    check(ldpc_b200_create(&c, &h_), "ldpc_b200_create");
    h_method_ = method;
}

int CLDPC_B200::run(int method) {
    ensure(method);
    check(ldpc_b200_decode(h_, fixInput, decodedBits, n_groups_, bf_iters.data(), its_per_group.data(), conv_iter.data()),
          "ldpc_b200_decode");
    return bf_iters[0];
}

void CLDPC_B200::Decode() { run(LDPC_B200_NMS); }
void CLDPC_B200::Decode_OMS() { run(LDPC_B200_OMS); }
void CLDPC_B200::Decode_FAID() { run(LDPC_B200_FAID_DTBF); }
int CLDPC_B200::Decode_OMSBF() { return run(LDPC_B200_OMS_BF); }
int CLDPC_B200::Decode_OMS_DTBF() { return run(LDPC_B200_OMS_DTBF); }
void CLDPC_B200::Decode_FAID_2B1C() { run(LDPC_B200_FAID_2B1C); }

int CLDPC_B200::DecodeDispatch(int m) {
    switch (m) {  // CSimulate.cpp:136-164
    case 1: Decode_OMS(); return -1;
    case 2: Decode_FAID(); return -1;
    case 3: return Decode_OMSBF();
    case 4: return Decode_OMS_DTBF();
    case 5: Decode_FAID_2B1C(); return -1;
    default: Decode(); return -1;
    }
}

void CLDPC_B200::GenMsgSeq() {
    for (size_t i = 0; i < (size_t)n_groups_ * 32 * LDPC_B200_K; ++i) inputBits[i] = rand() % 2;
}
void CLDPC_B200::Encode() { check(ldpc_b200_encode(h_, inputBits, outputBits, n_groups_), "ldpc_b200_encode"); }

void CLDPC_B200::FakeEncoder(const int* cw) {
    for (int g = 0; g < n_groups_; ++g) {
        int8_t* in = inputBits + (size_t)g * 32 * LDPC_B200_K;
        int8_t* out = outputBits + (size_t)g * 32 * LDPC_B200_N;
        for (int f = 0; f < 32; ++f) {
            for (int j = 0; j < LDPC_B200_K; ++j) in[f * LDPC_B200_K + j] = out[f * LDPC_B200_K + j] = (int8_t)(cw[j] > 0);
            for (int j = 0; j < LDPC_B200_M; ++j) out[32 * LDPC_B200_K + f * LDPC_B200_M + j] = (int8_t)(cw[LDPC_B200_K + j] > 0);
        }
    }
}

void CLDPC_B200::FakeEncoder() {
    std::vector<int> zero(LDPC_B200_N, 0);
    FakeEncoder(zero.data());
}

void CLDPC_B200::float2LimitChar_4bit(int8_t* output, const float* input, float scale, int length) {
    check(ldpc_b200_quantize(h_, input, output, length, scale), "ldpc_b200_quantize");
}

void CLDPC_B200::GenerateNoisyBlock(float Eb_N0, uint64_t seed, uint64_t first_frame_index) {
    check(ldpc_b200_generate(h_, outputBits, Eb_N0, seed, first_frame_index, n_groups_, nullptr, fixInput), "ldpc_b200_generate");
}

Statistic CLDPC_B200::CalculateErrors() {
    uint64_t c[LDPC_B200_NUM_COUNTERS] = {0};
    check(ldpc_b200_count_errors(h_, inputBits, decodedBits, n_groups_, c), "ldpc_b200_count_errors");
    Statistic s{};
    s.ErrorFrame = c[LDPC_B200_CNT_ERROR_FRAME];
    s.ErrorBits = c[LDPC_B200_CNT_ERROR_BITS];
    s.LT3ErrBitFrame = c[LDPC_B200_CNT_LT3_ERR_BIT_FRAME];
    return s;
}

Statistic CLDPC_B200::CalculateErrors(float*, int8_t*, int) { return CalculateErrors(); }
