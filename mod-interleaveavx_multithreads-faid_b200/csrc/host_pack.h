// host_pack.h -- host-side marshalling between the reference's byte-per-value buffers and the engine's packed layouts.
//
// The reference keeps one int8 per LLR (`fixInput`, 4-bit values in [-7,7], CLDPC.cpp:4524-4582) and one int8 per decoded
// bit (`decodedBits`, CLDPC.cpp:2268-2270).  Over PCIe that is 17 664 B per frame in each direction, and a pageable
// (malloc'ed, as in the reference) buffer has to be staged through pinned memory anyway.  The staging copy therefore
// packs on the way: two LLRs per byte towards the device (the native layout of ldpc_b200_decode_packed), one BIT per
// decoded bit back, expanded into the caller's int8 array by the same threads.  No decoding arithmetic runs here.
#pragma once
#include <cstddef>
#include <cstdint>

namespace ldpc {

struct HostPool;
// n_threads >= 1 (the caller is one of them).  cpus (optional, Linux cpu_set_t*, cpu_set_bytes long): the workers are pinned to
// that CPU set -- the NUMA node of the handle's GPU, so that pack / expand traffic stays on the socket the DMA engine uses.
HostPool* host_pool_create(int n_threads, const void* cpus = nullptr, size_t cpu_set_bytes = 0);
void host_pool_destroy(HostPool* p);
int host_pool_threads(const HostPool* p);

// fix: `groups` groups in the reference's two-region layout (per group: [32][K] info bytes, then [32][M] parity bytes).
// packed: [groups * 32][N / 2], low nibble = even code bit.  Returns false (output unspecified) if any value is outside
// [-8, 7] -- the caller then ships the bytes unpacked.
bool host_pack_llr(HostPool* p, const int8_t* fix, uint8_t* packed, int groups);
// hard: [frames][N / 32] words, bit n % 32 of word n / 32 = decoded bit n.  decoded: [frames][N] bytes 0 / 1.
void host_unpack_bits(HostPool* p, const uint32_t* hard, int8_t* decoded, int frames);

// Both at once, in ONE pass over the pool (one wake-up instead of two per chunk; the pack of chunk i and the expansion of the
// chunk that last used the same slot balance each other): either half may be empty (fix == nullptr / decoded == nullptr).
// Returns what host_pack_llr would.
bool host_stage_both(HostPool* p, const int8_t* fix, uint8_t* packed, int groups, const uint32_t* hard, int8_t* decoded, int frames);

// NUMA placement (Linux sysfs; every function degrades to "unknown" on other hosts).
//   numa_node_of_pci: node of the PCI device "dddd:bb:dd.f" (lower case), -1 if unknown
//   numa_node_cpus:   fills a cpu_set_t (cpu_set_bytes long) with the CPUs of the node that the calling thread may also run
//                     on; returns their number (0 if unknown)
int numa_node_of_pci(const char* bdf);
int numa_node_cpus(int node, void* cpus, size_t cpu_set_bytes);
// Runs fn() with the calling thread temporarily restricted to `cpus` (first-touch page placement of pinned allocations).
struct ScopedAffinity {
    ScopedAffinity(const void* cpus, size_t cpu_set_bytes);
    ~ScopedAffinity();
    unsigned char saved[128];
    bool active = false;
};

}  // namespace ldpc
