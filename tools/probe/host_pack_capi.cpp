// C entry points over csrc/host_pack.cpp for tests/test_host_pack.py (test infrastructure; the product calls the C++ API).
#include "host_pack.h"
#define LDPC_N_FOR_TEST 17664
extern "C" {
int hp_pack(const int8_t* fix, uint8_t* packed, int groups, int threads) {
    ldpc::HostPool* p = ldpc::host_pool_create(threads);
    const bool ok = ldpc::host_pack_llr(p, fix, packed, groups);
    ldpc::host_pool_destroy(p);
    return ok ? 1 : 0;
}
void hp_unpack(const uint32_t* hard, int8_t* decoded, int frames, int threads) {
    ldpc::HostPool* p = ldpc::host_pool_create(threads);
    ldpc::host_unpack_bits(p, hard, decoded, frames);
    ldpc::host_pool_destroy(p);
}
}

// Pool stress: `rounds` fused pack + expand passes on ONE pool, with pauses of `sleep_us` microseconds between some of them so
// that workers go through both the spinning and the blocked state; returns the number of rounds whose output was wrong.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>
extern "C" int hp_stress(const int8_t* fix, const uint8_t* packed_ref, const uint32_t* hard, const int8_t* decoded_ref, int groups,
                         int frames, int threads, int rounds, int sleep_us) {
    ldpc::HostPool* p = ldpc::host_pool_create(threads);
    std::vector<uint8_t> packed((size_t)groups * 32 * (LDPC_N_FOR_TEST / 2));
    std::vector<int8_t> decoded((size_t)frames * LDPC_N_FOR_TEST);
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
        std::memset(packed.data(), 0xEE, packed.size());
        std::memset(decoded.data(), 0x55, decoded.size());
        const bool ok = ldpc::host_stage_both(p, (r % 5 == 4) ? nullptr : fix, packed.data(), groups, hard, (r % 7 == 6) ? nullptr : decoded.data(), frames);
        if (!ok) ++bad;
        if (r % 5 != 4 && std::memcmp(packed.data(), packed_ref, packed.size()) != 0) ++bad;
        if (r % 7 != 6 && std::memcmp(decoded.data(), decoded_ref, decoded.size()) != 0) ++bad;
        if (sleep_us > 0 && r % 3 == 0) std::this_thread::sleep_for(std::chrono::microseconds(r % 2 ? sleep_us : sleep_us * 20));
    }
    ldpc::host_pool_destroy(p);
    return bad;
}

// Stand-alone throughput of the staging passes (no GPU activity): seconds per pass over `groups` groups, best of `rounds`, for
// pack alone, expand alone, the fused pass, and -- as the box's yardstick -- a plain streaming copy of the same number of bytes
// the fused pass moves, split over the same threads.  out[4] = {pack_s, expand_s, fused_s, copy_s}.
#include <immintrin.h>
#include <cstdlib>
__attribute__((target("avx512f"))) static void stream_copy512(int8_t* d, const int8_t* s, size_t n) {
    for (size_t i = 0; i < n; i += 64) _mm512_stream_si512((__m512i*)(d + i), _mm512_load_si512((const __m512i*)(s + i)));
}
static void stream_copy(int8_t* d, const int8_t* s, size_t n) {
    if (__builtin_cpu_supports("avx512f")) stream_copy512(d, s, n);
    else std::memcpy(d, s, n);
}
extern "C" int hp_bench(int groups, int threads, int rounds, double* out) {
    const size_t frames = (size_t)groups * 32;
    const size_t n_fix = frames * LDPC_N_FOR_TEST, n_packed = n_fix / 2, n_hard = frames * (LDPC_N_FOR_TEST / 32);
    int8_t* fix = (int8_t*)aligned_alloc(4096, n_fix);
    uint8_t* packed = (uint8_t*)aligned_alloc(4096, n_packed);
    uint32_t* hard = (uint32_t*)aligned_alloc(4096, n_hard * 4);
    int8_t* decoded = (int8_t*)aligned_alloc(4096, n_fix);
    if (!fix || !packed || !hard || !decoded) return -1;
    for (size_t i = 0; i < n_fix; ++i) fix[i] = (int8_t)((i * 2654435761u >> 13) % 15) - 7;
    for (size_t i = 0; i < n_hard; ++i) hard[i] = (uint32_t)(i * 2246822519u);
    std::memset(packed, 0, n_packed);
    std::memset(decoded, 0, n_fix);
    ldpc::HostPool* p = ldpc::host_pool_create(threads);
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int k = 0; k < 4; ++k) out[k] = 1e30;
    for (int r = 0; r < rounds + 1; ++r) {
        double t0 = now();
        ldpc::host_pack_llr(p, fix, packed, groups);
        double t1 = now();
        ldpc::host_unpack_bits(p, hard, decoded, (int)frames);
        double t2 = now();
        ldpc::host_stage_both(p, fix, packed, groups, hard, decoded, (int)frames);
        double t3 = now();
        if (r == 0) continue;
        out[0] = std::min(out[0], t1 - t0);
        out[1] = std::min(out[1], t2 - t1);
        out[2] = std::min(out[2], t3 - t2);
    }
    ldpc::host_pool_destroy(p);
    // yardstick: threads x (read a, stream-store b) over n_fix bytes  (n_fix read + n_fix written)
    {
        std::vector<std::thread> th;
        for (int r = 0; r < rounds + 1; ++r) {
            const double t0 = now();
            th.clear();
            for (int t = 0; t < threads; ++t)
                th.emplace_back([&, t] {
                    const size_t lo = n_fix / threads * t / 64 * 64, hi = (t == threads - 1) ? n_fix / 64 * 64 : n_fix / threads * (t + 1) / 64 * 64;
                    stream_copy(decoded + lo, fix + lo, hi - lo);
                });
            for (auto& x : th) x.join();
            const double t1 = now();
            if (r) out[3] = std::min(out[3], t1 - t0);
        }
    }
    free(fix); free(packed); free(hard); free(decoded);
    return 0;
}
