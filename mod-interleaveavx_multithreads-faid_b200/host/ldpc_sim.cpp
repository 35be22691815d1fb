// ldpc_sim -- Eb/N0 sweep with the reference's round structure, stop rule and result files (main.cpp:136-228), every
// round running entirely on the GPU through ldpc_b200_simulate (CSimulate::Run, CSimulate.cpp:92-180).
//
//   ldpc_sim [Profile.txt] [--groups-per-round G] [--seed S] [--max-frames F] [--device D] [--fixed-codeword]
//            [--out-dir DIR] [--first-frame N] [--collect]
//
// stdout: the Result.txt columns, tab separated, one line per Eb/N0 point (plus the average iteration count).
// --out-dir DIR additionally writes the reference's files into DIR with the reference's formats:
//   Result.txt      main.cpp:220-223   Eb/N0, TestFrame, ErrorFrame, ErrorBits, FER, BER, LT3ErrBitFrame, Time(s)
//   Temp.txt        main.cpp:194-207   running totals of the current point; the resume table holds the Philox checkpoint
//                                      (seed, next frame index) instead of the 3-LCG states of every thread
//   demod.txt       main.cpp:224-227   always zero in the reference (ModCalErr is commented out, CSimulate.cpp:129-131)
//   iterCount.txt   main.cpp:149-150, CSimulate.cpp:171-179   histogram of BF iterations per group
//   errorindex.txt / errorfloat.txt / errordecode.txt   CLDPC.cpp:4877-4991   error-frame dumps once FER < 1e-5
//                                      (collectflag, main.cpp:190-192) or with --collect
// Error-frame dumps: the rounds run fused on the device and only counters come back; Philox is counter based, so a round
// that reported errors is REPLAYED step by step through the C-ABI (gen_msg_seq -> encode -> generate -> demap -> decode)
// and the failing frames are written out.  With several GPUs one process per GPU is started by the caller and
// counters are merged with ldpc_b200_allreduce_counters.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "ldpc_b200.h"

namespace {
constexpr int N = LDPC_B200_N, K = LDPC_B200_K, M = LDPC_B200_M, Z = 256;

#define TRY(expr)                                                 \
    do {                                                          \
        if (expr) {                                               \
            fprintf(stderr, "%s\n", ldpc_b200_last_error());      \
            return 1;                                             \
        }                                                         \
    } while (0)

// Replays groups [g0, g0 + n) of the stream (global group index) and appends the reference's three dump files for every
// frame with info-bit errors.  Returns the number of frames dumped, -1 on error.
int replay_and_dump(ldpc_b200_handle* h, const ldpc_b200_config& cfg, const int8_t* fixed_cw, float snr, uint64_t seed,
                    uint64_t g0, int n, const std::string& dir) {
    const int reuse = cfg.codeword_reuse > 0 ? cfg.codeword_reuse : 50;
    const int nsym_f = cfg.mod_type == 1 ? 32 * N : 2 * 32 * N / cfg.mod_type;
    std::vector<int8_t> info(32 * K), tx(32 * N), fix(32 * N), dec(32 * N);
    std::vector<float> sym(nsym_f), llr(32 * N);
    int dumped = 0;
    for (int g = 0; g < n; ++g) {
        const uint64_t G = g0 + g;
        if (fixed_cw) {
            for (int f = 0; f < 32; ++f) {
                memcpy(&info[(size_t)f * K], fixed_cw, K);
                memcpy(&tx[(size_t)f * K], fixed_cw, K);
                memcpy(&tx[(size_t)32 * K + (size_t)f * M], fixed_cw + K, M);
            }
        } else {
            if (ldpc_b200_gen_msg_seq(h, seed, (G / reuse) * reuse * 32, 1, info.data())) return -1;
            if (ldpc_b200_encode(h, info.data(), tx.data(), 1)) return -1;
        }
        if (ldpc_b200_generate(h, tx.data(), snr, seed, G * 32, 1, sym.data(), fix.data())) return -1;
        if (cfg.mod_type != 1) {
            if (ldpc_b200_demap(h, sym.data(), 1, llr.data(), fix.data())) return -1;
        } else {
            llr = sym;  // BPSK: the received amplitudes are the LLRs (CSimulate.cpp:121-124)
        }
        if (ldpc_b200_decode(h, fix.data(), dec.data(), 1, nullptr, nullptr, nullptr)) return -1;
        for (int f = 0; f < 32; ++f) {
            std::vector<int> eb;
            for (int j = 0; j < K; ++j)
                if (dec[(size_t)f * N + j] != info[(size_t)f * K + j]) eb.push_back(j);
            if (eb.empty()) continue;
            std::vector<int> ec;  // parity-bit mismatches are located but not counted (CLDPC.cpp:4858-4868)
            for (int j = 0; j < M; ++j)
                if (dec[(size_t)f * N + K + j] != tx[(size_t)32 * K + (size_t)f * M + j]) ec.push_back(j);
            std::ofstream eout(dir + "/errorindex.txt", std::ios::app), nout(dir + "/errorfloat.txt", std::ios::app),
                dout(dir + "/errordecode.txt", std::ios::app);
            eout << "ErrorFrame: " << f << std::endl;
            eout << "ErrorBit Num: " << eb.size() << std::endl;
            eout << "Errorbit Block: ";
            for (int j : eb) eout << j / Z << "\t";
            eout << std::endl << "Errobit Index: ";
            for (int j : eb) eout << j % Z << "\t";
            eout << std::endl << "Errorcheck Num: " << ec.size() << std::endl << "Errorcheck Block: ";
            for (int j : ec) eout << j / Z << "\t";
            eout << std::endl << "Errorcheck Index: ";
            for (int j : ec) eout << j % Z << "\t";
            eout << std::endl;
            nout << "ErrorFloat=[ ";
            for (int j = 0; j < K; ++j) nout << llr[(size_t)f * K + j] << "\t";
            for (int j = 0; j < M; ++j) nout << llr[(size_t)32 * K + (size_t)f * M + j] << "\t";
            nout << "];" << std::endl << "ErrorChar=[";
            for (int j = 0; j < K; ++j) nout << (int)fix[(size_t)f * K + j] << "\t";
            for (int j = 0; j < M; ++j) nout << (int)fix[(size_t)32 * K + (size_t)f * M + j] << "\t";
            nout << "];" << std::endl << std::endl;
            dout << "Decodedbits=[";
            for (int j = 0; j < N; ++j) dout << (int)dec[(size_t)f * N + j] << "\t";
            dout << "];" << std::endl << "inputbits=[";
            for (int j = 0; j < K; ++j) dout << (int)info[(size_t)f * K + j] << "\t";
            dout << "];" << std::endl << "outputbits=[";
            for (int j = 0; j < K; ++j) dout << (int)tx[(size_t)f * K + j] << "\t";
            for (int j = 0; j < M; ++j) dout << (int)tx[(size_t)32 * K + (size_t)f * M + j] << "\t";
            dout << "];" << std::endl << std::endl;
            ++dumped;
        }
    }
    return dumped;
}
}  // namespace

int main(int argc, char** argv) {
    const char* profile = "Profile.txt";
    int groups = 50;  // the reference runs 50 blocks of 32 frames per thread and round (CSimulate.cpp:117)
    uint64_t seed = 101, max_frames = 0, frame0 = 0;
    int device = 0;
    bool fixed = false, force_collect = false;
    std::string dir;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--groups-per-round") && i + 1 < argc) groups = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--max-frames") && i + 1 < argc) max_frames = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--first-frame") && i + 1 < argc) frame0 = strtoull(argv[++i], nullptr, 10) / 32 * 32;
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--out-dir") && i + 1 < argc) dir = argv[++i];
        else if (!strcmp(argv[i], "--fixed-codeword")) fixed = true;
        else if (!strcmp(argv[i], "--collect")) force_collect = true;
        else profile = argv[i];
    }
    ldpc_b200_config cfg;
    memset(&cfg, 0, sizeof cfg);
    TRY(ldpc_b200_read_profile(profile, &cfg, -1));
    cfg.device = device;
    ldpc_b200_handle* h = nullptr;
    TRY(ldpc_b200_create(&cfg, &h));
    static int8_t zero_cw[N];
    printf("Eb/N0\tTestFrame\tErrorFrame\tErrorBits\tFER\tBER\tLT3ErrBitFrame\tTime(s)\tavgIter\n");
    int collectflag = force_collect ? 1 : 0;  // the reference's global, set once FER < 1e-5 (main.cpp:190-192)
    using std::setw;
    for (float snr = cfg.snr_start; snr < cfg.snr_end; snr += cfg.snr_pass) {  // main.cpp:136
        uint64_t c[LDPC_B200_NUM_COUNTERS] = {0};
        if (!dir.empty()) {
            for (const char* fn : {"errorindex.txt", "errorfloat.txt", "errordecode.txt", "iterCount.txt"}) {
                std::ofstream o(dir + "/" + fn, std::ios::app);  // main.cpp:149-153
                o << "Eb/N0: " << setw(5) << snr << "scale=" << cfg.scale << std::endl;
            }
        }
        auto t0 = std::chrono::steady_clock::now();
        double FER = 1, BER = 1;
        while (c[LDPC_B200_CNT_TEST_FRAME] < 1000 || c[LDPC_B200_CNT_ERROR_FRAME] < 20) {  // main.cpp:164
            const uint64_t err_before = c[LDPC_B200_CNT_ERROR_FRAME];
            TRY(ldpc_b200_simulate(h, fixed ? zero_cw : nullptr, snr, seed, frame0, groups, c));
            const uint64_t tf = c[LDPC_B200_CNT_TEST_FRAME], ef = c[LDPC_B200_CNT_ERROR_FRAME], ebits = c[LDPC_B200_CNT_ERROR_BITS];
            BER = (double)(ebits > 0 ? ebits : 1) / ((double)tf * K);  // "assume one is wrong" (main.cpp:186-188)
            FER = (double)(ef > 0 ? ef : 1) / (double)tf;
            if (FER < 1e-5) collectflag = 1;
            if (!dir.empty() && collectflag && ef > err_before) {
                const int d = replay_and_dump(h, cfg, fixed ? zero_cw : nullptr, snr, seed, frame0 / 32, groups, dir);
                if (d < 0 || (uint64_t)d != ef - err_before) {
                    fprintf(stderr, "replay of groups [%llu, +%d) found %d error frames, the round counted %llu: %s\n",
                            (unsigned long long)(frame0 / 32), groups, d, (unsigned long long)(ef - err_before), ldpc_b200_last_error());
                    return 1;
                }
            }
            frame0 += (uint64_t)groups * 32;
            if (!dir.empty()) {
                std::ofstream tout(dir + "/Temp.txt", std::ios::out);  // main.cpp:194-207
                tout << setw(5) << snr << '\t' << setw(20) << tf << '\t' << setw(15) << ef << '\t' << setw(20) << ebits << '\t'
                     << setw(20) << FER << '\t' << setw(20) << BER << '\t' << setw(15) << c[LDPC_B200_CNT_LT3_ERR_BIT_FRAME] << '\t' << std::endl;
                // resume information: Philox is counter based, (seed, next frame index) replaces lastSeed[threads][3]
                tout << "const unsigned long long lastPhilox[2] = {" << seed << "," << frame0 << "};\n";
            }
            if (tf > 1000 && ef > 20) break;  // main.cpp:209-211
            if (max_frames && tf >= max_frames) break;
        }
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double tf = (double)c[LDPC_B200_CNT_TEST_FRAME];
        printf("%.2f\t%llu\t%llu\t%llu\t%.3e\t%.3e\t%llu\t%.2f\t%.2f\n", snr, (unsigned long long)c[0], (unsigned long long)c[1],
               (unsigned long long)c[2], c[1] / tf, c[2] / (tf * K), (unsigned long long)c[3], t,
               (double)c[LDPC_B200_CNT_MS_ITERS_SUM] / (double)c[LDPC_B200_CNT_GROUPS]);
        fflush(stdout);
        if (!dir.empty()) {
            std::ofstream fout(dir + "/Result.txt", std::ios::app), demod(dir + "/demod.txt", std::ios::app),
                iter(dir + "/iterCount.txt", std::ios::app);
            fout << setw(5) << snr << '\t' << setw(20) << c[0] << '\t' << setw(15) << c[1] << '\t' << setw(20) << c[2] << '\t'
                 << setw(20) << FER << '\t' << setw(20) << BER << '\t' << setw(15) << c[3] << '\t' << setw(15) << t << '\t' << std::endl;
            demod << setw(5) << snr << '\t' << setw(20) << 0.0 << '\t' << setw(20) << 0.0 << '\t' << setw(20) << 0.0 << '\t' << std::endl;
            for (int i = 1; i <= 50; ++i)  // CSimulate.cpp:171-179 (methods 3 and 4 count BF iterations per group)
                if (c[LDPC_B200_CNT_BF_HIST + i]) iter << i << ": " << c[LDPC_B200_CNT_BF_HIST + i] << std::endl;
        }
    }
    ldpc_b200_destroy(h);
    return 0;
}
