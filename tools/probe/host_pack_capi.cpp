// C entry points over csrc/host_pack.cpp for tests/test_host_pack.py (test infrastructure; the product calls the C++ API).
#include "host_pack.h"
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
