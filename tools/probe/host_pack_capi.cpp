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
