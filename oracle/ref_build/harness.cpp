// C-ABI harness around the UNMODIFIED reference translation units (compiled from /root/reference by
// oracle/Makefile into oracle/_ref/).  TEST INFRASTRUCTURE ONLY: it is the checker the CUDA path and the
// plain-C restatement (oracle/ldpc_oracle.c) are validated against, and the optional CPU baseline of
// bench.py (--impl reference / cpu_baseline).  Nothing in the product path may link or load this.
//
// It mirrors CSimulate::Initial/Configure/Run (CSimulate.cpp:41-59,67-75,103-166) because CSimulate.cpp
// itself does not compile as shipped (stray tokens at CSimulate.cpp:123).
#include "CChannel.h"
#include "CLDPC.h"
#include "CModulate.h"
#include "CTool.h"
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <unistd.h>
#include <vector>

int collectflag = 0;  // main.cpp:14
int MAX_THREADS = 0;  // main.cpp:15
extern int CodeWord_sym[_NoVar];  // Codeword.h:4 (non-const global, defined in CLDPC.cpp's TU)

#ifdef REF_INSTRUMENTED
// Only the instrumented build (generated copies of the decoder TUs, see oracle/Makefile) defines these.
int g_ref_iter_enter = 0;             // number of evaluations of `while (nombre_iterations--)`
unsigned char g_ref_errsum_log[64 * 32];  // per-iteration per-lane error_sum as seen right after the early-stop test
int g_ref_errsum_n = 0;
#endif

struct RefSim {
    CLDPC* ldpc;
    CModulate* mod;
    CChannel* ch;
    int mod_type;
};

static int dispatch(CLDPC* l, int method) {
    int bf = -1;
    switch (method) {  // CSimulate.cpp:136-164
    case 0: l->Decode(); break;
    case 1: l->Decode_OMS(); break;
    case 2: l->Decode_FAID(); break;
    case 3: bf = l->Decode_OMSBF(); break;
    case 4: bf = l->Decode_OMS_DTBF(); break;
    case 5: l->Decode_FAID_2B1C(); break;
    case 101: l->Decode1(); break;  // never dispatched by CSimulate (CSimulate.cpp:204 comment only): generic NMS init, test use
    default: l->Decode(); break;
    }
    return bf;
}

static CLDPC* new_ldpc(int max_iter) {
    CLDPC* l = new CLDPC();
    l->Initial(32, max_iter);
    // CLDPC.cpp:4803-4804 mallocs VN_weight_ without zero-fill before counting (UB).  The intended
    // semantics are the true column weights: recount on a zeroed array (public member + public method).
    memset(l->VN_weight_, 0, _NoVar);
    l->VN_weight_count();
    return l;
}

extern "C" {

int ref_variant_instrumented() {
#ifdef REF_INSTRUMENTED
    return 1;
#else
    return 0;
#endif
}

// Profile.txt is re-read from the cwd inside every Decode*/CalculateErrors (CLDPC.cpp:216-217).
int ref_write_profile(const char* dir, float snr_start, float snr_pass, float snr_end, int method, int max_iter,
                      int mod_type, int interleave, int f1, int f2, float scale) {
    char path[4096];
    snprintf(path, sizeof path, "%s/Profile.txt", dir);
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f,
            "Simulation parameter\nStartSNR: %g\nSNRPass: %g\nEndSNR: %g\nDecodeMethod: %d\nMaxIteration: %d\n"
            "Modulation Parameter:\nmodType: %d\nInterleaveModType: %d\nNMS  Factor:\nFactor_1: %d\nFactor_2: %d\n"
            "noFrames: 32\nscale: %.9g\nMatrix Factor\nFileName: 50GPON-CP12\nZ: 256\n",
            snr_start, snr_pass, snr_end, method, max_iter, mod_type, interleave, f1, f2, scale);
    fclose(f);
    return 0;
}

int ref_chdir(const char* dir) { return chdir(dir); }

void* ref_ldpc_create(int max_iter) { return new_ldpc(max_iter); }
void ref_ldpc_destroy(void* h) { delete (CLDPC*)h; }
void ref_ldpc_set_iterations(void* h, int max_iter) { ((CLDPC*)h)->nb_iteration = max_iter; }

// fixInput: int8[32*N] in the reference layout; decodedBits: int8[32*N] frame-major 0/1.
int ref_decode(void* h, int method, const int8_t* fixInput, int8_t* decodedBits) {
    CLDPC* l = (CLDPC*)h;
    memcpy(l->fixInput, fixInput, 32 * _NoVar);
#ifdef REF_INSTRUMENTED
    g_ref_iter_enter = 0;
    g_ref_errsum_n = 0;
#endif
    int bf = dispatch(l, method);
    memcpy(decodedBits, l->decodedBits, 32 * _NoVar);
    return bf;
}

// Instrumented build only: iterations executed by the last ref_decode and the per-lane error_sum log.
int ref_last_iterations(unsigned char* errsum_log, int max_entries) {
#ifdef REF_INSTRUMENTED
    int n = g_ref_errsum_n < max_entries ? g_ref_errsum_n : max_entries;
    if (errsum_log) memcpy(errsum_log, g_ref_errsum_log, (size_t)n * 32);
    return g_ref_iter_enter - 1;
#else
    (void)errsum_log; (void)max_entries;
    return -1;
#endif
}

void ref_quantize_4bit(void* h, int8_t* out, const float* in, float scale, int length) {
    ((CLDPC*)h)->float2LimitChar_4bit(out, in, scale, length);
}

int ref_quantize_bits(void* h, int8_t* out, const float* in, float scale, int length, int bits) {
    CLDPC* l = (CLDPC*)h;
    switch (bits) {
    case 6: l->float2LimitChar_6bit(out, in, scale, length); break;
    case 5: l->float2LimitChar_5bit(out, in, scale, length); break;
    case 4: l->float2LimitChar_4bit(out, in, scale, length); break;
    case 3: l->float2LimitChar_3bit(out, in, scale, length); break;
    case 2: l->float2LimitChar_2bit(out, in, scale, length); break;
    case 1: l->float2LimitChar_1bit(out, in, scale, length); break;
    default: return -1;
    }
    return 0;
}

void ref_transpose(const int8_t* src, int8_t* dst, int n) {
    uchar_transpose_avx((__m256i*)src, (__m256i*)dst, n);
}
void ref_itranspose(const int8_t* src, int8_t* dst, int n) {
    uchar_itranspose_avx((__m256i*)src, (__m256i*)dst, n);
}

void ref_vn_weight(void* h, int8_t* out) { memcpy(out, ((CLDPC*)h)->VN_weight_, _NoVar); }

// ---- full chain: mirror of CSimulate::Initial / Run -------------------------------------------
void* ref_sim_create(int max_iter, int mod_type, int interleave, int seed) {
    RefSim* s = new RefSim;
    s->ldpc = new_ldpc(max_iter);
    s->mod = new CModulate();
    s->ch = new CChannel();
    s->mod_type = mod_type;
    s->mod->ModulationType = mod_type;
    s->mod->InterleaveModType = interleave;
    s->mod->Initial(32UL * _NoVar);
    s->ch->RandomSeed = seed;
    s->ch->Initial(s->mod->SymbolLen, 0);
    return s;
}
void ref_sim_destroy(void* h) {
    RefSim* s = (RefSim*)h;
    delete s->ldpc; delete s->mod; delete s->ch; delete s;
}
void* ref_sim_ldpc(void* h) { return ((RefSim*)h)->ldpc; }
double ref_sim_rate(void* h) { return ((RefSim*)h)->ldpc->m_Rate; }

// FakeEncoder on a caller-supplied codeword (same codeword in all 32 lanes), then interleave + map.
// Outputs (optional): inputBits int8[32*K], outputBits int8[32*N] (two-region layout), ModSeq complex64[32*N/m].
void ref_sim_set_codeword(void* h, const int8_t* codeword, int8_t* inputBits, int8_t* outputBits, float* modseq) {
    RefSim* s = (RefSim*)h;
    for (int i = 0; i < _NoVar; ++i) CodeWord_sym[i] = codeword[i];
    s->ldpc->FakeEncoder();
    s->mod->BeforeModulationInterleaver(s->ldpc->outputBits);
    s->mod->Modulation(s->mod->InterLeaveSeq);
    if (inputBits) memcpy(inputBits, s->ldpc->inputBits, 32 * NmoinsK);
    if (outputBits) memcpy(outputBits, s->ldpc->outputBits, 32 * _NoVar);
    if (modseq) memcpy(modseq, s->mod->ModSeq, sizeof(MKL_Complex8) * s->mod->SymbolLen);
}

// Arbitrary per-frame transmitted bits (two-region outputBits layout) instead of FakeEncoder.
void ref_sim_set_output_bits(void* h, const int8_t* inputBits, const int8_t* outputBits) {
    RefSim* s = (RefSim*)h;
    memcpy(s->ldpc->inputBits, inputBits, 32 * NmoinsK);
    memcpy(s->ldpc->outputBits, outputBits, 32 * _NoVar);
    s->mod->BeforeModulationInterleaver(s->ldpc->outputBits);
    s->mod->Modulation(s->mod->InterLeaveSeq);
}

// One noise block: AWGN(sigma/sqrt(2)) -> demap -> de-interleave -> 4-bit quantise  (CSimulate.cpp:126-132).
// sigma is the value CSimulate::Configure computes (CSimulate.cpp:70-74).
void ref_sim_noise_block(void* h, float sigma, float scale, float* symbols, float* demod, float* deint,
                         int8_t* fixInput) {
    RefSim* s = (RefSim*)h;
    s->ch->AWGNChannel(s->mod->ModSeq, sigma / sqrt(2));
    s->mod->Demodulation(s->ch->SymbolSeq);
    s->mod->AfterDeModulationDeInterleaver();
    s->ldpc->float2LimitChar_4bit(s->ldpc->fixInput, s->mod->DeInterLeaveSeq, scale, BitsOverChannel * 32);
    if (symbols) memcpy(symbols, s->ch->SymbolSeq, sizeof(MKL_Complex8) * s->mod->SymbolLen);
    if (demod) memcpy(demod, s->mod->DemodSeq, sizeof(float) * 32 * _NoVar);
    if (deint) memcpy(deint, s->mod->DeInterLeaveSeq, sizeof(float) * 32 * _NoVar);
    if (fixInput) memcpy(fixInput, s->ldpc->fixInput, 32 * _NoVar);
}

// Demap + de-interleave + quantise caller-supplied noisy symbols (for producer parity on identical symbols).
void ref_sim_demap_block(void* h, const float* symbols, float scale, float* demod, float* deint, int8_t* fixInput) {
    RefSim* s = (RefSim*)h;
    memcpy(s->ch->SymbolSeq, symbols, sizeof(MKL_Complex8) * s->mod->SymbolLen);
    s->mod->Demodulation(s->ch->SymbolSeq);
    s->mod->AfterDeModulationDeInterleaver();
    s->ldpc->float2LimitChar_4bit(s->ldpc->fixInput, s->mod->DeInterLeaveSeq, scale, BitsOverChannel * 32);
    if (demod) memcpy(demod, s->mod->DemodSeq, sizeof(float) * 32 * _NoVar);
    if (deint) memcpy(deint, s->mod->DeInterLeaveSeq, sizeof(float) * 32 * _NoVar);
    if (fixInput) memcpy(fixInput, s->ldpc->fixInput, 32 * _NoVar);
}

// BPSK (mod_type 1 objects only): CModulate::BPSKModulation on caller-supplied outputBits, and the receive side of
// CSimulate.cpp:121-124 on caller-supplied noisy amplitudes (the MKL Gaussian stream itself is stubbed, mkl.h).
void ref_sim_bpsk_modulate(void* h, const int8_t* outputBits, float* modseq) {
    RefSim* s = (RefSim*)h;
    memcpy(s->ldpc->outputBits, outputBits, 32 * _NoVar);
    s->mod->BPSKModulation(s->ldpc->outputBits);
    memcpy(modseq, s->mod->BPSKModSeq, sizeof(float) * s->mod->SymbolLen);
}
void ref_sim_bpsk_receive(void* h, const float* noisy, float scale, int8_t* fixInput) {
    RefSim* s = (RefSim*)h;
    memcpy(s->ch->BPSKSymbol, noisy, sizeof(float) * s->mod->SymbolLen);
    s->ldpc->float2LimitChar_4bit(s->ldpc->fixInput, s->ch->BPSKSymbol, scale, BitsOverChannel * 32);
    memcpy(fixInput, s->ldpc->fixInput, 32 * _NoVar);
}

void ref_sim_rng_state(void* h, unsigned long* st) {
    RefSim* s = (RefSim*)h;
    st[0] = s->ch->RS.IX; st[1] = s->ch->RS.IY; st[2] = s->ch->RS.IZ;
}

// Decode the block currently in ldpc->fixInput and score it with CalculateErrors (CSimulate.cpp:136-169).
int ref_sim_decode_and_count(void* h, int method, int8_t* decodedBits, unsigned long* stats3) {
    RefSim* s = (RefSim*)h;
    int bf = dispatch(s->ldpc, method);
    Statistic t = s->ldpc->CalculateErrors(s->mod->DeInterLeaveSeq, s->ldpc->fixInput, 0);
    stats3[0] = t.ErrorFrame; stats3[1] = t.ErrorBits; stats3[2] = t.LT3ErrBitFrame;
    if (decodedBits) memcpy(decodedBits, s->ldpc->decodedBits, 32 * _NoVar);
    return bf;
}

// CalculateErrors on caller-supplied buffers (two-region inputBits/outputBits, frame-major decodedBits).
void ref_calc_errors(void* h, const int8_t* inputBits, const int8_t* outputBits, const int8_t* decodedBits,
                     unsigned long* stats3) {
    CLDPC* l = (CLDPC*)h;
    memcpy(l->inputBits, inputBits, 32 * NmoinsK);
    memcpy(l->outputBits, outputBits, 32 * _NoVar);
    memcpy(l->decodedBits, decodedBits, 32 * _NoVar);
    Statistic t = l->CalculateErrors(nullptr, l->fixInput, 0);
    stats3[0] = t.ErrorFrame; stats3[1] = t.ErrorBits; stats3[2] = t.LT3ErrBitFrame;
}

// ---- CPU baseline: T threads, each with a private CLDPC, looping Decode*() over the given groups ----
// Returns frames decoded per second (aggregate); "as shipped" = includes the per-call Profile.txt re-read and
// the transposes inside Decode*().  groups: int8[n_groups][32*N] in the reference layout.
double ref_bench_decode(int method, int max_iter, int n_threads, double min_seconds, const int8_t* groups,
                        int n_groups, long* frames_done) {
    std::atomic<long> total(0);
    std::atomic<int> go(0);
    std::vector<std::thread> th;
    std::vector<CLDPC*> objs(n_threads);
    for (int t = 0; t < n_threads; ++t) objs[t] = new_ldpc(max_iter);
    auto worker = [&](int t) {
        CLDPC* l = objs[t];
        while (!go.load()) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        long n = 0;
        int g = t % n_groups;
        for (;;) {
            memcpy(l->fixInput, groups + (size_t)g * 32 * _NoVar, 32 * _NoVar);
            dispatch(l, method);
            n += 32;
            g = (g + 1) % n_groups;
            double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el >= min_seconds) break;
        }
        total += n;
    };
    for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
    auto t0 = std::chrono::steady_clock::now();
    go.store(1);
    for (auto& x : th) x.join();
    double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto p : objs) delete p;
    if (frames_done) *frames_done = total.load();
    return total.load() / el;
}

}  // extern "C"
