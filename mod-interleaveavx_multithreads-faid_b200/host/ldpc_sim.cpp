// ldpc_sim -- Eb/N0 sweep with the reference's round structure and stop rule (main.cpp:136-228), every round running
// entirely on the GPU through ldpc_b200_simulate (CSimulate::Run, CSimulate.cpp:92-180).
//
//   ldpc_sim [Profile.txt] [--groups-per-round G] [--seed S] [--max-frames F] [--device D] [--fixed-codeword]
//
// Writes the Result.txt columns of the reference (main.cpp:216-223) to stdout; with several GPUs one process per GPU
// is started by the caller and counters are merged with ldpc_b200_allreduce_counters.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ldpc_b200.h"

int main(int argc, char** argv) {
    const char* profile = "Profile.txt";
    int groups = 50;  // the reference runs 50 blocks of 32 frames per thread and round (CSimulate.cpp:117)
    uint64_t seed = 101, max_frames = 0;
    int device = 0;
    bool fixed = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--groups-per-round") && i + 1 < argc) groups = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--max-frames") && i + 1 < argc) max_frames = strtoull(argv[++i], nullptr, 10);
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--fixed-codeword")) fixed = true;
        else profile = argv[i];
    }
    ldpc_b200_config cfg;
    memset(&cfg, 0, sizeof cfg);
    if (ldpc_b200_read_profile(profile, &cfg, -1)) {
        fprintf(stderr, "%s\n", ldpc_b200_last_error());
        return 1;
    }
    cfg.device = device;
    ldpc_b200_handle* h = nullptr;
    if (ldpc_b200_create(&cfg, &h)) {
        fprintf(stderr, "%s\n", ldpc_b200_last_error());
        return 1;
    }
    static int8_t zero_cw[LDPC_B200_N];
    printf("Eb/N0\tTestFrame\tErrorFrame\tErrorBits\tFER\tBER\tLT3ErrBitFrame\tTime(s)\tavgIter\n");
    uint64_t frame0 = 0;
    for (float snr = cfg.snr_start; snr < cfg.snr_end; snr += cfg.snr_pass) {  // main.cpp:136
        uint64_t c[LDPC_B200_NUM_COUNTERS] = {0};
        auto t0 = std::chrono::steady_clock::now();
        while (c[LDPC_B200_CNT_TEST_FRAME] < 1000 || c[LDPC_B200_CNT_ERROR_FRAME] < 20) {  // main.cpp:164
            if (ldpc_b200_simulate(h, fixed ? zero_cw : nullptr, snr, seed, frame0, groups, c)) {
                fprintf(stderr, "%s\n", ldpc_b200_last_error());
                return 1;
            }
            frame0 += (uint64_t)groups * 32;
            if (max_frames && c[LDPC_B200_CNT_TEST_FRAME] >= max_frames) break;
        }
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double tf = (double)c[LDPC_B200_CNT_TEST_FRAME];
        printf("%.2f\t%llu\t%llu\t%llu\t%.3e\t%.3e\t%llu\t%.2f\t%.2f\n", snr, (unsigned long long)c[0], (unsigned long long)c[1],
               (unsigned long long)c[2], c[1] / tf, c[2] / (tf * LDPC_B200_K), (unsigned long long)c[3], t,
               (double)c[LDPC_B200_CNT_MS_ITERS_SUM] / (double)c[LDPC_B200_CNT_GROUPS]);
        fflush(stdout);
    }
    ldpc_b200_destroy(h);
    return 0;
}
