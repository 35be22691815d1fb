// CSimulate::Run (CSimulate.cpp:92-180) written against the CLDPC-shaped shim: FakeEncoder on a codeword read from a
// file of '0'/'1' characters, n blocks of (noise -> decode -> CalculateErrors).  Prints the counters; used by
// tests/test_gpu_host_shim.py to check the C++ surface against the Python/C-ABI path on the same Philox stream.
//   run_like_csimulate <Profile.txt> <codeword.txt> <Eb/N0> <seed> <blocks>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "CLDPC_b200.h"

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    try {
        Parameter_Simulation p;
        ReadProfile(&p, argv[1]);
        std::ifstream f(argv[2]);
        std::vector<int> cw;
        char ch;
        while (f >> ch) cw.push_back(ch == '1');
        if ((int)cw.size() != LDPC_B200_N) return 3;
        const float ebn0 = (float)atof(argv[3]);
        const uint64_t seed = strtoull(argv[4], nullptr, 10);
        const int blocks = atoi(argv[5]);
        CLDPC_B200 ldpc;
        ldpc.Initial(p, 1, 0, -1);
        ldpc.FakeEncoder(cw.data());
        unsigned long TestFrame = 0, ErrorFrame = 0, ErrorBits = 0, LT3 = 0;
        int BFiters_[51] = {0};
        for (int i = 0; i < blocks; ++i) {  // CSimulate.cpp:117
            TestFrame += 32;
            ldpc.GenerateNoisyBlock(ebn0, seed, (uint64_t)i * 32);
            const int bf = ldpc.DecodeDispatch(p.decode_method);
            if (bf >= 0) BFiters_[bf]++;
            const Statistic t = ldpc.CalculateErrors();
            ErrorFrame += t.ErrorFrame;
            ErrorBits += t.ErrorBits;
            LT3 += t.LT3ErrBitFrame;
        }
        printf("%lu %lu %lu %lu\n", TestFrame, ErrorFrame, ErrorBits, LT3);
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
