// dropin_csimulate.cpp -- the reference's OWN front end (CTool, CChannel, CModulate, CLDPC quantiser / CalculateErrors:
// compiled unmodified from /root/reference by oracle/Makefile) driving three decoders through the body of CSimulate::Run
// (CSimulate.cpp:41-59 Initial, :67-75 Configure, :103-169 Run; CSimulate.cpp itself does not compile as shipped,
// stray tokens at :123, so its body is restated here with FAKE_ENCODE 1):
//
//   ref    the reference's CLDPC::Decode*()                                                   (the baseline run)
//   cabi   INTEGRATION.md section 3, first patch: the `switch (decode_method)` of CSimulate.cpp:136-164 replaced by
//          ONE call ldpc_b200_decode(gpu, ldpc->fixInput, ldpc->decodedBits, 1, &BFiter, &its, nullptr); everything else,
//          including ldpc->CalculateErrors, is the reference's
//   shim   `CLDPC` swapped for the CLDPC-shaped class CLDPC_B200 (host/CLDPC_b200.h): FakeEncoder, float2LimitChar_4bit,
//          Decode*(), CalculateErrors all go to the GPU engine; CChannel / CModulate stay the reference's
//
// Every arm owns its object set and starts the reference's 3-LCG channel from the same seed (seed[0] = 101,
// CSimulate.cpp:11), so all arms see the same noise.  Per block it prints TestFrame-independent observables:
//   <arm> <block> <ErrorFrame> <ErrorBits> <LT3ErrBitFrame> <BFiter> <fnv1a64(decodedBits)>
// and exits non-zero if any arm differs from `ref` in any of them.  TEST INFRASTRUCTURE (links the reference's objects).
//
//   dropin_csimulate <method> <Eb/N0> <blocks> <codeword.txt | zero> [arms=ref,cabi,shim]      (reads ./Profile.txt)
#include "CChannel.h"
#include "CLDPC.h"
#include "CModulate.h"
#include "CTool.h"
// after the reference's headers: reuses their Parameter_Simulation / Statistic
#include "CLDPC_b200.h"
#include "ldpc_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

int collectflag = 0;  // main.cpp:14
int MAX_THREADS = 0;  // main.cpp:15
extern int CodeWord_sym[_NoVar];  // Codeword.h:4

static const int kSeed0 = 101;  // seed[0], CSimulate.cpp:11

struct Row {
    unsigned long ef, eb, lt3;
    int bf;
    unsigned long long hash;
    bool operator==(const Row& o) const { return ef == o.ef && eb == o.eb && lt3 == o.lt3 && bf == o.bf && hash == o.hash; }
};

static unsigned long long fnv(const int8_t* p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    return h;
}

static float sigma_of(const Parameter_Simulation& p, double rate, float snr) {  // CSimulate.cpp:67-75
    if (p.mod_type == 1) return (float)(1.0 / sqrt(2.0 * rate * p.mod_type * pow(10.0, 0.1 * snr)));
    return (float)(1.0 / sqrt(rate * p.mod_type * pow(10.0, 0.1 * snr)));
}

// CSimulate::Initial for the two reference-owned front-end objects (CSimulate.cpp:41-59)
static void init_front_end(const Parameter_Simulation& p, CModulate*& modulate, CChannel*& channel) {
    channel = new CChannel();
    modulate = new CModulate();
    modulate->ModulationType = p.mod_type;
    modulate->InterleaveModType = p.interleavemod_type;
    modulate->Initial(32UL * _NoVar);  // m_frame * (m_N - m_PunLen - m_ShortenLen), both 0 as shipped
    channel->RandomSeed = kSeed0;
    channel->Initial(modulate->SymbolLen, 0);
}

static int dispatch_ref(CLDPC* ldpc, int decode_method) {  // CSimulate.cpp:136-164
    int BFiter = -1;
    switch (decode_method) {
    case 0: ldpc->Decode(); break;
    case 1: ldpc->Decode_OMS(); break;
    case 2: ldpc->Decode_FAID(); break;
    case 3: BFiter = ldpc->Decode_OMSBF(); break;
    case 4: BFiter = ldpc->Decode_OMS_DTBF(); break;
    case 5: ldpc->Decode_FAID_2B1C(); break;
    default: ldpc->Decode(); break;
    }
    return BFiter;
}

// arms `ref` and `cabi`: every object is the reference's; only the decode dispatch differs
static std::vector<Row> run_reference_objects(const Parameter_Simulation& p, float snr, int blocks, bool use_gpu) {
    CLDPC* ldpc = new CLDPC();
    CModulate* modulate;
    CChannel* channel;
    ldpc->Initial(p.nb_frames, p.Max_Iteration);
    memset(ldpc->VN_weight_, 0, _NoVar);  // CLDPC.cpp:4803-4804 counts into un-zeroed malloc memory (UB): recount on zeros
    ldpc->VN_weight_count();
    init_front_end(p, modulate, channel);
    const float sigma = sigma_of(p, ldpc->m_Rate, snr), scale = p.scale;
    ldpc_b200_handle* gpu = nullptr;
    if (use_gpu) {  // what a maintainer adds to CSimulate::Initial
        ldpc_b200_config cfg;
        ldpc_b200_default_config(&cfg, p.decode_method, -1);
        cfg.max_iteration = p.Max_Iteration; cfg.mod_type = p.mod_type; cfg.interleave_mod_type = p.interleavemod_type;
        cfg.factor_1 = p.Factor_1; cfg.factor_2 = p.Factor_2; cfg.scale = p.scale;
        cfg.device = 0;
        if (ldpc_b200_create(&cfg, &gpu)) { fprintf(stderr, "%s\n", ldpc_b200_last_error()); exit(EXIT_FAILURE); }
    }
    std::vector<Row> rows;
    ldpc->FakeEncoder();
    if (modulate->ModulationType == 1) {
        modulate->BPSKModulation(ldpc->outputBits);
    } else {
        modulate->BeforeModulationInterleaver(ldpc->outputBits);
        modulate->Modulation(modulate->InterLeaveSeq);
    }
    for (int i = 0; i < blocks; ++i) {
        if (modulate->ModulationType == 1) {
            channel->BPSKAWGNChannel(modulate->BPSKModSeq, sigma);  // MKL stream is stubbed (mkl.h): noiseless, arms still agree
            ldpc->float2LimitChar_4bit(ldpc->fixInput, channel->BPSKSymbol, scale, BitsOverChannel * 32);
        } else {
            channel->AWGNChannel(modulate->ModSeq, sigma / sqrt(2));
            modulate->Demodulation(channel->SymbolSeq);
            modulate->AfterDeModulationDeInterleaver();
            ldpc->float2LimitChar_4bit(ldpc->fixInput, modulate->DeInterLeaveSeq, scale, BitsOverChannel * 32);
        }
        int BFiter = -1;
        if (!use_gpu) {
            BFiter = dispatch_ref(ldpc, p.decode_method);
        } else {
            int its;
            if (ldpc_b200_decode(gpu, ldpc->fixInput, ldpc->decodedBits, /*n_groups=*/1, &BFiter, &its, nullptr)) {
                fprintf(stderr, "%s\n", ldpc_b200_last_error());
                exit(EXIT_FAILURE);
            }
            if (p.decode_method != 3 && p.decode_method != 4) BFiter = -1;  // only the BF variants return it in the reference
        }
        Statistic Test = ldpc->CalculateErrors(modulate->DeInterLeaveSeq, ldpc->fixInput, collectflag);
        rows.push_back({Test.ErrorFrame, Test.ErrorBits, Test.LT3ErrBitFrame, BFiter, fnv(ldpc->decodedBits, 32 * _NoVar)});
    }
    if (gpu) ldpc_b200_destroy(gpu);
    delete ldpc; delete modulate; delete channel;
    return rows;
}

// arm `shim`: the same body with `CLDPC` replaced by `CLDPC_B200`
static std::vector<Row> run_shim(const Parameter_Simulation& p, float snr, int blocks) {
    CLDPC_B200* ldpc = new CLDPC_B200();
    CModulate* modulate;
    CChannel* channel;
    ldpc->Initial(p.nb_frames, p.Max_Iteration);
    init_front_end(p, modulate, channel);
    const float sigma = sigma_of(p, ldpc->m_Rate, snr), scale = p.scale;
    std::vector<Row> rows;
    ldpc->FakeEncoder(CodeWord_sym);
    if (modulate->ModulationType == 1) {
        modulate->BPSKModulation(ldpc->outputBits);
    } else {
        modulate->BeforeModulationInterleaver(ldpc->outputBits);
        modulate->Modulation(modulate->InterLeaveSeq);
    }
    for (int i = 0; i < blocks; ++i) {
        if (modulate->ModulationType == 1) {
            channel->BPSKAWGNChannel(modulate->BPSKModSeq, sigma);
            ldpc->float2LimitChar_4bit(ldpc->fixInput, channel->BPSKSymbol, scale, BitsOverChannel * 32);
        } else {
            channel->AWGNChannel(modulate->ModSeq, sigma / sqrt(2));
            modulate->Demodulation(channel->SymbolSeq);
            modulate->AfterDeModulationDeInterleaver();
            ldpc->float2LimitChar_4bit(ldpc->fixInput, modulate->DeInterLeaveSeq, scale, BitsOverChannel * 32);
        }
        int BFiter = -1;
        switch (p.decode_method) {  // CSimulate.cpp:136-164, verbatim structure
        case 0: ldpc->Decode(); break;
        case 1: ldpc->Decode_OMS(); break;
        case 2: ldpc->Decode_FAID(); break;
        case 3: BFiter = ldpc->Decode_OMSBF(); break;
        case 4: BFiter = ldpc->Decode_OMS_DTBF(); break;
        case 5: ldpc->Decode_FAID_2B1C(); break;
        default: ldpc->Decode(); break;
        }
        Statistic Test = ldpc->CalculateErrors(modulate->DeInterLeaveSeq, ldpc->fixInput, collectflag);
        rows.push_back({Test.ErrorFrame, Test.ErrorBits, Test.LT3ErrBitFrame, BFiter, fnv(ldpc->decodedBits, 32 * _NoVar)});
    }
    delete ldpc; delete modulate; delete channel;
    return rows;
}

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: dropin_csimulate <method> <Eb/N0> <blocks> <codeword.txt|zero> [arms]\n");
        return 2;
    }
    Parameter_Simulation p;
    ReadProfile(&p);  // the reference's own parser, ./Profile.txt
    p.decode_method = atoi(argv[1]);
    const float snr = (float)atof(argv[2]);
    const int blocks = atoi(argv[3]);
    if (strcmp(argv[4], "zero") != 0) {
        std::ifstream f(argv[4]);
        char ch;
        int n = 0;
        while (n < _NoVar && f >> ch) CodeWord_sym[n++] = ch == '1';
        if (n != _NoVar) { fprintf(stderr, "codeword file must hold %d characters 0/1\n", _NoVar); return 2; }
    }
    const std::string arms = argc > 5 ? argv[5] : "ref,cabi,shim";
    // The reference's decoders re-read Factor_1/2 from ./Profile.txt, whose DecodeMethod line may differ from argv[1]; only
    // the factors, MaxIteration, modType, scale of the file matter to them.
    std::vector<Row> ref, other;
    int bad = 0;
    try {
        if (arms.find("ref") != std::string::npos) {
            ref = run_reference_objects(p, snr, blocks, false);
            for (int i = 0; i < blocks; ++i) printf("ref %d %lu %lu %lu %d %016llx\n", i, ref[i].ef, ref[i].eb, ref[i].lt3, ref[i].bf, ref[i].hash);
        }
        for (const char* arm : {"cabi", "shim"}) {
            if (arms.find(arm) == std::string::npos) continue;
            other = strcmp(arm, "cabi") == 0 ? run_reference_objects(p, snr, blocks, true) : run_shim(p, snr, blocks);
            for (int i = 0; i < blocks; ++i) {
                printf("%s %d %lu %lu %lu %d %016llx\n", arm, i, other[i].ef, other[i].eb, other[i].lt3, other[i].bf, other[i].hash);
                if (!ref.empty() && !(other[i] == ref[i])) ++bad;
            }
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    if (bad) { fprintf(stderr, "%d block(s) differ from the reference decoder\n", bad); return 3; }
    return 0;
}
