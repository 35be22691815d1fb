#include "host_pack.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main(int argc, char** argv) {
    const int N = 17664, K = 14592, M = 3072, HW = N / 32;
    int groups = 256, threads = argc > 1 ? atoi(argv[1]) : 8;
    std::vector<int8_t> fix((size_t)groups * 32 * N), dec((size_t)groups * 32 * N + 64);
    std::vector<uint8_t> packed((size_t)groups * 32 * N / 2);
    std::vector<uint32_t> hard((size_t)groups * 32 * HW);
    srand(1);
    for (auto& v : fix) v = (int8_t)(rand() % 15 - 7);
    for (auto& v : hard) v = (uint32_t)rand() * 2654435761u;
    auto* p = ldpc::host_pool_create(threads);
    bool ok = ldpc::host_pack_llr(p, fix.data(), packed.data(), groups);
    // verify
    size_t errs = 0;
    for (int f = 0; f < groups * 32; ++f)
        for (int n = 0; n < N; ++n) {
            int g = f >> 5, fg = f & 31;
            int8_t v = n < K ? fix[(size_t)g * 32 * N + (size_t)fg * K + n] : fix[(size_t)g * 32 * N + 32 * (size_t)K + (size_t)fg * M + n - K];
            int nib = (packed[(size_t)f * N / 2 + n / 2] >> (4 * (n & 1))) & 15;
            if (((nib ^ 8) - 8) != v) ++errs;
        }
    int8_t* d = dec.data();
    ldpc::host_unpack_bits(p, hard.data(), d, groups * 32);
    for (size_t f = 0; f < (size_t)groups * 32; ++f)
        for (int n = 0; n < N; ++n)
            if (d[f * N + n] != (int8_t)((hard[f * HW + n / 32] >> (n % 32)) & 1)) ++errs;
    fix[12345] = 9;
    bool ok2 = ldpc::host_pack_llr(p, fix.data(), packed.data(), groups);
    fix[12345] = 0;
    printf("ok=%d errs=%zu range-detect=%d\n", ok, errs, !ok2);
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) ldpc::host_pack_llr(p, fix.data(), packed.data(), groups);
        auto t1 = std::chrono::steady_clock::now();
        for (int i = 0; i < 5; ++i) ldpc::host_unpack_bits(p, hard.data(), d, groups * 32);
        auto t2 = std::chrono::steady_clock::now();
        double fr = 5.0 * groups * 32;
        printf("threads %d: pack %.2f Mframes/s, unpack %.2f Mframes/s\n", threads, fr / std::chrono::duration<double>(t1 - t0).count() / 1e6,
               fr / std::chrono::duration<double>(t2 - t1).count() / 1e6);
    }
    ldpc::host_pool_destroy(p);
}
