// host_pack.cpp -- see host_pack.h.  Compiled by g++ (not nvcc): AVX-512 bodies selected at run time, scalar otherwise.
#include "host_pack.h"

#include <immintrin.h>
#include <x86intrin.h>
#include <pthread.h>
#include <sched.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "ldpc_code_tables.h"

namespace ldpc {

namespace {
constexpr int kN = LDPC_N, kK = LDPC_K, kM = LDPC_M, kHW = LDPC_N / 32;
}

// Persistent workers; run(n, fn) executes fn(i) for i in [0, n) on all of them (dynamic chunks of 8) and returns when every
// index is done.  The calling thread takes part.  Chunks of a host-buffer decode arrive every millisecond or so, so a worker
// that finds no work spins on the generation counter for a short while (a condition-variable wake-up of 15 threads costs
// 50-100 us, several per cent of a chunk) and only then blocks.
struct HostPool {
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go;
    std::atomic<uint64_t> generation{0};
    std::atomic<int> pending{0};
    std::atomic<int> sleepers{0};
    std::atomic<bool> stop{false};
    int n_items = 0;
    const std::function<void(int)>* fn = nullptr;
    std::atomic<int> next{0};

    void work() {
        // dynamic chunks of 8 items: frames are uniform, but threads are not (other load on the host)
        for (;;) {
            const int i0 = next.fetch_add(8, std::memory_order_relaxed);
            if (i0 >= n_items) break;
            const int i1 = i0 + 8 < n_items ? i0 + 8 : n_items;
            for (int i = i0; i < i1; ++i) (*fn)(i);
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            // spin for ~60 us of TSC time (the gap between the two passes of a chunk, and between chunks when the GPU keeps
            // up), then block
            bool got = false;
            const unsigned long long t0 = __rdtsc();
            for (;;) {
                if (stop.load(std::memory_order_acquire)) return;
                if (generation.load(std::memory_order_acquire) != seen) { got = true; break; }
                if (__rdtsc() - t0 > 150000ull) break;
                _mm_pause();
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                sleepers.fetch_add(1, std::memory_order_seq_cst);
                cv_go.wait(lk, [&] { return stop.load(std::memory_order_acquire) || generation.load(std::memory_order_acquire) != seen; });
                sleepers.fetch_sub(1, std::memory_order_seq_cst);
                if (stop.load(std::memory_order_acquire)) return;
            }
            seen = generation.load(std::memory_order_acquire);
            work();
            pending.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    void run(int n, const std::function<void(int)>& f) {
        if (n <= 0) return;
        if (workers.empty() || n < 16) {
            for (int i = 0; i < n; ++i) f(i);
            return;
        }
        fn = &f;
        n_items = n;
        next.store(0, std::memory_order_relaxed);
        pending.store((int)workers.size(), std::memory_order_relaxed);
        {
            // the generation bump is published under the mutex so that a worker about to block cannot miss it
            std::lock_guard<std::mutex> lk(mu);
            generation.fetch_add(1, std::memory_order_seq_cst);
        }
        if (sleepers.load(std::memory_order_seq_cst) > 0) cv_go.notify_all();
        work();
        while (pending.load(std::memory_order_acquire) != 0) _mm_pause();  // every worker has seen this generation and drained it
    }
};

HostPool* host_pool_create(int n_threads, const void* cpus, size_t cpu_set_bytes) {
    HostPool* p = new HostPool();
    for (int i = 1; i < n_threads; ++i) {
        p->workers.emplace_back([p] { p->loop(); });
        if (cpus && cpu_set_bytes) pthread_setaffinity_np(p->workers.back().native_handle(), cpu_set_bytes, (const cpu_set_t*)cpus);
    }
    return p;
}

int numa_node_of_pci(const char* bdf) {
    char path[256];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

int numa_node_cpus(int node, void* cpus, size_t cpu_set_bytes) {
    if (node < 0 || !cpus || cpu_set_bytes < sizeof(cpu_set_t)) return 0;
    char path[256];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = fopen(path, "r");
    if (!f) return 0;
    char buf[4096];
    const bool ok = fgets(buf, sizeof buf, f) != nullptr;
    fclose(f);
    if (!ok) return 0;
    cpu_set_t allowed, *out = (cpu_set_t*)cpus;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return 0;
    CPU_ZERO(out);
    int n = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {  // "0-15,32-47"
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, out); ++n; }
    }
    return n;
}

ScopedAffinity::ScopedAffinity(const void* cpus, size_t cpu_set_bytes) {
    static_assert(sizeof(cpu_set_t) <= sizeof(saved), "cpu_set_t larger than the save area");
    if (!cpus || cpu_set_bytes < sizeof(cpu_set_t)) return;
    if (sched_getaffinity(0, sizeof(cpu_set_t), (cpu_set_t*)saved) != 0) return;
    if (sched_setaffinity(0, sizeof(cpu_set_t), (const cpu_set_t*)cpus) != 0) return;
    active = true;
}
ScopedAffinity::~ScopedAffinity() {
    if (active) sched_setaffinity(0, sizeof(cpu_set_t), (const cpu_set_t*)saved);
}
void host_pool_destroy(HostPool* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->mu);
        p->stop.store(true, std::memory_order_seq_cst);
    }
    p->cv_go.notify_all();
    for (auto& t : p->workers) t.join();
    delete p;
}
int host_pool_threads(const HostPool* p) { return p ? (int)p->workers.size() + 1 : 0; }

namespace {

bool pack_streaming();

// n is a multiple of 128 (K = 114 * 128, M = 24 * 128).  Returns true if every value is in [-8, 7].
__attribute__((target("avx512f,avx512bw"))) bool pack_row_avx512(const int8_t* src, uint8_t* dst, int n) {
    const __m512i lo_mask = _mm512_set1_epi16(0x000F), hi_mask = _mm512_set1_epi16(0x00F0), eight = _mm512_set1_epi8(8);
    const __m512i fifteen = _mm512_set1_epi8(15);
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 63) == 0 && pack_streaming();
    __mmask64 bad = 0;
    for (int i = 0; i < n; i += 128) {
        const __m512i x0 = _mm512_loadu_si512(src + i), x1 = _mm512_loadu_si512(src + i + 64);
        bad |= _mm512_cmpgt_epu8_mask(_mm512_add_epi8(x0, eight), fifteen) | _mm512_cmpgt_epu8_mask(_mm512_add_epi8(x1, eight), fifteen);
        // 16-bit lane = (odd byte << 8) | even byte  ->  low byte (even & 15) | (odd & 15) << 4
        const __m512i y0 = _mm512_or_si512(_mm512_and_si512(x0, lo_mask), _mm512_and_si512(_mm512_srli_epi16(x0, 4), hi_mask));
        const __m512i y1 = _mm512_or_si512(_mm512_and_si512(x1, lo_mask), _mm512_and_si512(_mm512_srli_epi16(x1, 4), hi_mask));
        const __m512i v = _mm512_inserti64x4(_mm512_castsi256_si512(_mm512_cvtepi16_epi8(y0)), _mm512_cvtepi16_epi8(y1), 1);
        // the staging buffer is read next by the DMA engine, not by this core: streaming store, no read-for-ownership
        if (aligned) _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i / 2), v);
        else _mm512_storeu_si512(dst + i / 2, v);
    }
    return bad == 0;
}
bool pack_row_scalar(const int8_t* src, uint8_t* dst, int n) {
    int bad = 0;
    for (int i = 0; i < n; i += 2) {
        const int a = src[i], b = src[i + 1];
        bad |= (a < -8) | (a > 7) | (b < -8) | (b > 7);
        dst[i / 2] = (uint8_t)((a & 15) | ((b & 15) << 4));
    }
    return bad == 0;
}

// one frame: 552 words -> 17 664 bytes (276 blocks of 64).  Streaming stores need 64-byte aligned addresses; malloc / numpy arrays
// are 16-byte aligned (the alignment the C-ABI asks for), and a frame is a multiple of 64 bytes, so every frame of such an array
// starts `head` = 16, 32 or 48 bytes before a line boundary: those bytes and the matching tail go out as masked stores, the 275
// lines between them as streaming stores whose 64 mask bits are read at a byte offset into the packed words.  (With ordinary
// stores the destination lines are read before they are written: 17.7 KB more host-memory traffic per frame, 32 instead of 42
// Gbit/s for the host-buffer call on such arrays, profiles/r02_e2e_pageable.log.)
__attribute__((target("avx512f,avx512bw"))) void unpack_frame_avx512(const uint32_t* hard, int8_t* dst) {
    const __m512i one = _mm512_set1_epi8(1);
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(dst) & 63);
    if (mis == 0) {
        for (int w = 0; w < kHW; w += 2) {
            const __mmask64 m = (uint64_t)hard[w] | ((uint64_t)hard[w + 1] << 32);
            _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + 32 * w), _mm512_maskz_mov_epi8(m, one));  // the caller reads it later, not now
        }
        return;
    }
    if (mis & 7) {  // not even 8-byte aligned: the mask bits would not start on a byte of the packed words
        for (int w = 0; w < kHW; w += 2) {
            const __mmask64 m = (uint64_t)hard[w] | ((uint64_t)hard[w + 1] << 32);
            _mm512_storeu_si512(dst + 32 * w, _mm512_maskz_mov_epi8(m, one));
        }
        return;
    }
    const unsigned head = 64 - mis;  // bytes (= code bits) before the first line boundary, a multiple of 8
    const uint8_t* bits = reinterpret_cast<const uint8_t*>(hard);  // little endian: bit n of the frame is bit n % 8 of byte n / 8
    uint64_t m0;
    std::memcpy(&m0, bits, 8);
    _mm512_mask_storeu_epi8(dst, ((uint64_t)1 << head) - 1, _mm512_maskz_mov_epi8(m0, one));
    const int lines = (kN - (int)head) / 64;  // whole lines after the head
    for (int l = 0; l < lines; ++l) {
        uint64_t m;
        std::memcpy(&m, bits + (head + 64u * l) / 8, 8);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + head + 64 * l), _mm512_maskz_mov_epi8(m, one));
    }
    const int done = (int)head + 64 * lines, tail = kN - done;  // tail = mis bytes
    if (tail > 0) {
        uint64_t m = 0;
        std::memcpy(&m, bits + done / 8, (size_t)tail / 8);
        _mm512_mask_storeu_epi8(dst + done, ((uint64_t)1 << tail) - 1, _mm512_maskz_mov_epi8(m, one));
    }
}
void unpack_frame_scalar(const uint32_t* hard, int8_t* dst) {
    for (int w = 0; w < kHW; ++w) {
        const uint32_t x = hard[w];
        for (int b = 0; b < 32; ++b) dst[32 * w + b] = (int8_t)((x >> b) & 1u);
    }
}

// LDPC_B200_PACK_NT=0: the nibble-packed staging buffer is written with ordinary stores (it may then stay in the last-level
// cache, where the copy engine's reads can hit, instead of going to DRAM and back); default 1 = streaming stores
bool pack_streaming() {
    static const bool v = [] { const char* e = getenv("LDPC_B200_PACK_NT"); return !e || atoi(e) != 0; }();
    return v;
}

bool have_avx512() {
    static const bool v = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    return v;
}

}  // namespace

bool host_pack_llr(HostPool* p, const int8_t* fix, uint8_t* packed, int groups) {
    static_assert(kK % 128 == 0 && kM % 128 == 0 && kHW % 2 == 0, "row lengths must be multiples of the vector width");
    const bool fast = have_avx512();
    std::atomic<int> bad{0};
    const std::function<void(int)> body = [&](int f) {
        const int g = f >> 5, fg = f & 31;
        const int8_t* info = fix + (size_t)g * 32 * kN + (size_t)fg * kK;
        const int8_t* par = fix + (size_t)g * 32 * kN + (size_t)32 * kK + (size_t)fg * kM;
        uint8_t* dst = packed + (size_t)f * (kN / 2);
        bool ok = fast ? pack_row_avx512(info, dst, kK) : pack_row_scalar(info, dst, kK);
        ok = (fast ? pack_row_avx512(par, dst + kK / 2, kM) : pack_row_scalar(par, dst + kK / 2, kM)) && ok;
        if (!ok) bad.store(1, std::memory_order_relaxed);
    };
    p->run(groups * 32, body);
    if (fast) _mm_sfence();
    return bad.load() == 0;
}

bool host_stage_both(HostPool* p, const int8_t* fix, uint8_t* packed, int groups, const uint32_t* hard, int8_t* decoded, int frames) {
    const bool fast = have_avx512();
    const int n_pack = fix ? groups * 32 : 0, n_unpack = decoded ? frames : 0;
    std::atomic<int> bad{0};
    const std::function<void(int)> body = [&](int i) {
        if (i < n_pack) {
            const int f = i, g = f >> 5, fg = f & 31;
            const int8_t* info = fix + (size_t)g * 32 * kN + (size_t)fg * kK;
            const int8_t* par = fix + (size_t)g * 32 * kN + (size_t)32 * kK + (size_t)fg * kM;
            uint8_t* dst = packed + (size_t)f * (kN / 2);
            bool ok = fast ? pack_row_avx512(info, dst, kK) : pack_row_scalar(info, dst, kK);
            ok = (fast ? pack_row_avx512(par, dst + kK / 2, kM) : pack_row_scalar(par, dst + kK / 2, kM)) && ok;
            if (!ok) bad.store(1, std::memory_order_relaxed);
        } else {
            const int f = i - n_pack;
            if (fast) unpack_frame_avx512(hard + (size_t)f * kHW, decoded + (size_t)f * kN);
            else unpack_frame_scalar(hard + (size_t)f * kHW, decoded + (size_t)f * kN);
        }
    };
    p->run(n_pack + n_unpack, body);
    if (fast) _mm_sfence();
    return bad.load() == 0;
}

void host_unpack_bits(HostPool* p, const uint32_t* hard, int8_t* decoded, int frames) {
    const bool fast = have_avx512();
    const std::function<void(int)> body = [&](int f) {
        if (fast) unpack_frame_avx512(hard + (size_t)f * kHW, decoded + (size_t)f * kN);
        else unpack_frame_scalar(hard + (size_t)f * kHW, decoded + (size_t)f * kN);
    };
    p->run(frames, body);
    if (fast) _mm_sfence();
}

}  // namespace ldpc
