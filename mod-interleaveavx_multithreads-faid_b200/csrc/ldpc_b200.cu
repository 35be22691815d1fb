// ldpc_b200.cu -- C-ABI (include/ldpc_b200.h) over the sm_100a kernels.  Host side only orchestrates:
// configuration, chunking, pinned staging, streams, CUDA-event timing.  There is NO CPU decode path: every
// entry point that computes fails with LDPC_B200_ENODEV when no CUDA device is usable.
#include "ldpc_b200.h"

#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "decode_kernels.cuh"
#include "code_tables_dev.cuh"
#include "bf_kernels.cuh"
#include "decode_launch.h"
#include "frame_kernels.cuh"
#include "host_pack.h"
#include "host_params.h"

using namespace ldpc;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(_e == cudaErrorMemoryAllocation ? LDPC_B200_ENOMEM : LDPC_B200_ECUDA,            \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));                             \
    } while (0)

// CDecoder_FAID.cpp:12-127, CDecoder_FAID_2B1C.cpp:12-47 (all four weight-class rows are equal in the reference)
const int8_t kLutSets[4][6][8] = {
    {{0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 4, 4, 4, 4}, {0, 1, 1, 3, 3, 4, 4, 4}, {0, 1, 1, 3, 3, 3, 6, 6}, {0, 1, 1, 3, 3, 3, 7, 7}},
    {{0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 4, 4, 4, 4}, {1, 1, 1, 1, 4, 4, 4, 4}, {1, 1, 1, 1, 5, 5, 5, 5}, {1, 1, 1, 1, 6, 6, 6, 6}},
    {{0, 0, 2, 2, 2, 2, 2, 2}, {0, 0, 2, 2, 2, 2, 2, 2}, {1, 1, 1, 3, 3, 3, 3, 3}, {1, 1, 1, 4, 4, 4, 4, 4}, {1, 1, 1, 5, 5, 5, 5, 5}, {1, 1, 1, 6, 6, 6, 6, 6}},
    {{0, 0, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 2, 3, 3, 3, 3}, {0, 1, 1, 3, 3, 4, 4, 4}, {0, 1, 1, 3, 3, 3, 6, 6}, {0, 1, 1, 3, 3, 3, 7, 7}},
};
const int8_t kLutEf[8] = {2, 3, 3, 4, 5, 6, 6, 7};  // CDecoder_FAID.cpp:130-165

void method_constants(ldpc_b200_config* c, int method, int lut_variant) {
    if (lut_variant < 0 || lut_variant > 3) lut_variant = (method == LDPC_B200_FAID_2B1C) ? LDPC_B200_LUT_HYBRID : LDPC_B200_LUT_FAID3;
    for (int it = 0; it < 6; ++it)
        for (int w = 0; w < 4; ++w)
            for (int a = 0; a < 8; ++a) {
                c->v2c_lut[it][w][a] = kLutSets[lut_variant][it][a];
                c->v2c_lut_ef[it][w][a] = kLutEf[a];
            }
    const bool m5 = method == LDPC_B200_FAID_2B1C;
    c->ef_elimination = m5 ? 1 : 0;          // CDecoder_FAID.cpp:5 / CDecoder_FAID_2B1C.cpp:6
    c->ef_floor_err_count = m5 ? 50 : 0;     // CDecoder_FAID.cpp:193 / CDecoder_FAID_2B1C.cpp:117
    c->ef_floor_iter_thresh = m5 ? 6 : -1;   // CDecoder_FAID.cpp:194 / CDecoder_FAID_2B1C.cpp:118
    c->oms_floor_err_count = 100;            // CDecoder_OMS.cpp:26
    c->oms_floor_iter_thresh = 4;            // CDecoder_OMS.cpp:27
    c->oms_mode = 1;                         // CDecoder_OMS.cpp:3
    c->oms_offset = 1;                       // CDecoder_OMS.cpp:6
    c->regular_col_weight = 3;               // CTool.h:6
    c->hard2_threshold = 13;                 // CDecoder_FAID_2B1C.cpp:6130
    c->dtbf_delta = 1;
    c->dtbf_alpha = 1;
    c->dtbf_L0 = c->dtbf_L1 = 0;
    switch (method) {
    case LDPC_B200_FAID_DTBF: c->bf_mode = LDPC_B200_BF_DTBF; c->bf_max_iter = 10; c->dtbf_L0 = 50; break;   // CDecoder_FAID.cpp:167-170,208
    case LDPC_B200_OMS_BF: c->bf_mode = LDPC_B200_BF_PLAIN; c->bf_max_iter = 50; break;                       // CDecoder_OMSBF.cpp:30
    case LDPC_B200_OMS_DTBF: c->bf_mode = LDPC_B200_BF_DTBF; c->bf_max_iter = 50; c->dtbf_L1 = 50; break;     // CDecoder_OMS_DTBF.cpp:6-9,35
    case LDPC_B200_FAID_2B1C: c->bf_mode = LDPC_B200_BF_2B1C; c->bf_max_iter = 10; c->dtbf_L0 = 100; break;   // CDecoder_FAID_2B1C.cpp:87-90,128
    default: c->bf_mode = LDPC_B200_BF_NONE; c->bf_max_iter = 0; break;
    }
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_mid = nullptr, ev_k1 = nullptr, ev_done = nullptr;
    bool timing_pending = false;
    int8_t* d_in = nullptr;       // staged input chunk (reference layout or packed)
    int8_t* d_out = nullptr;      // staged output chunk
    uint32_t* final_hard = nullptr;
    uint32_t* snap = nullptr;
    uint32_t* grp_cnt = nullptr;
    int32_t* first_zero = nullptr;
    unsigned int* work_counter = nullptr;  // ticket counter of the persistent decode kernel
    int32_t *d_bf = nullptr, *d_its = nullptr, *d_conv = nullptr;
    int32_t *h_bf = nullptr, *h_its = nullptr, *h_conv = nullptr;  // pinned
    // host staging (host_pack.h): pinned packed mirrors of the chunk, allocated on first use
    uint8_t* h_in_packed = nullptr;
    uint32_t* h_out_packed = nullptr;
    int8_t* unpack_dst = nullptr;  // caller's decodedBits of the chunk whose packed decisions are (about to be) in h_out_packed
    int unpack_frames = 0;
    // caller's per-group outputs of the chunk in flight (copied out of the pinned mirrors when the chunk has drained)
    int32_t *bf_dst = nullptr, *its_dst = nullptr, *conv_dst = nullptr;
    int info_groups = 0;
};

}  // namespace

struct ldpc_b200_handle {
    ldpc_b200_config cfg;
    int kind = KIND_NMS;
    int planes = 1;
    bool has_syndrome = false;
    int chunk_groups = 0;
    bool chunk_default = false;  // config left chunk_groups at 0
    std::vector<Slot> slots;
    std::vector<Slot> dslots;    // extra slots of the hybrid host-buffer path (chunks copied as they are), allocated on first use
    float last_kernel_ms = 0.f;
    float last_decode_ms = 0.f, last_finalize_ms = 0.f;
    int last_launches = 0;
    size_t fin_smem = 0;
    unsigned long long* d_dbg = nullptr;  // [2] bounds-check record of LDPC_DEBUG_BOUNDS builds (ldpc_b200_debug_bounds)
    int fin_wpf = kHW, fin_unsat_bufs = 1;  // finalize_kernel's shared-memory layout (fin_layout_words)
    int max_iteration_alloc = 0;  // MaxIteration the scratch (snapshots, group counters) was sized for
    // host staging threads (nullptr = the caller's buffers go over PCIe as they are)
    HostPool* pool = nullptr;
    bool stage_out = false, stage_in = false;
    uint64_t last_h2d_bytes = 0, last_d2h_bytes = 0;
    int last_direct_chunks = 0, last_staged_chunks = 0;
    // LDPC_B200_HOST_REGISTER=1: pageable caller arrays are page-locked with cudaHostRegister on first use and remembered by
    // address, so that later calls on the same buffers (the reference decodes out of ONE fixInput / decodedBits pair for the
    // whole run, CLDPC.h:123-124) are copied by the copy engines directly
    struct Registered { const void* ptr; size_t bytes; };
    std::vector<Registered> registered;  // hybrid host-buffer path: how the last call's chunks were routed
    // NUMA placement: CPUs of the GPU's node that this process may use (empty set = unknown / disabled with LDPC_B200_NUMA=0)
    cpu_set_t numa_cpus;
    int numa_node = -1, numa_ncpu = 0;
    // frame-generation state
    FrameState fs;
};

namespace {

int upload_tables(const ldpc_b200_config& c) {
    // once per device (the tables never change); serialised so that handles created concurrently from several host threads
    // (the reference's model: one object set per pthread) neither race on the flag nor rewrite a symbol a kernel is reading
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lock(mu);
    if (c.device >= 0 && c.device < 64 && done[c.device]) return LDPC_B200_OK;
    CodeTables ct;
    memcpy(ct.circ_col, ldpc_circ_col, sizeof ct.circ_col);
    memcpy(ct.circ_shift, ldpc_circ_shift, sizeof ct.circ_shift);
    memcpy(ct.col_layer, ldpc_col_layer, sizeof ct.col_layer);
    memcpy(ct.col_lshift, ldpc_col_lshift, sizeof ct.col_lshift);
    memcpy(ct.layer_start, ldpc_layer_start, sizeof ct.layer_start);
    memcpy(ct.col_start, ldpc_col_start, sizeof ct.col_start);
    memcpy(ct.col_weight, ldpc_col_weight, sizeof ct.col_weight);
    memcpy(ct.hpinv, ldpc_hpinv, sizeof ct.hpinv);
    CUDA_TRY(cudaMemcpyToSymbol(c_code, &ct, sizeof ct));
    if (c.device >= 0 && c.device < 64) done[c.device] = true;
    return LDPC_B200_OK;
}

int validate(const ldpc_b200_config& c) {
    if (c.struct_size != sizeof(ldpc_b200_config)) return fail(LDPC_B200_EINVAL, "config.struct_size mismatch");
    if (c.abi_version != LDPC_B200_ABI_VERSION) return fail(LDPC_B200_EINVAL, "config.abi_version mismatch");
    if (c.nb_frames != 32) return fail(LDPC_B200_EINVAL, "noFrames must be 32 (one __m256i of byte lanes in the reference)");
    if (c.Z != 256) return fail(LDPC_B200_EINVAL, "Z must be 256 (50G-PON code)");
    if (c.max_iteration < 0 || c.max_iteration > kMaxIterCap) return fail(LDPC_B200_EINVAL, "MaxIteration must be in [0, 1000]");
    if (!(c.mod_type == 1 || c.mod_type == 2 || c.mod_type == 4 || c.mod_type == 6 || c.mod_type == 8))
        return fail(LDPC_B200_EINVAL, "modType must be 1, 2, 4, 6 or 8 (CModulate.cpp:64-92)");
    if (c.oms_mode < 0 || c.oms_mode > 1 || c.oms_offset < 0 || c.oms_offset > 7) return fail(LDPC_B200_EINVAL, "oms_mode must be 0 or 1, oms_offset in [0,7]");
    if (c.codeword_reuse < 0) return fail(LDPC_B200_EINVAL, "codeword_reuse must be >= 0");
    if (c.quant_bits < 0 || c.quant_bits > 6) return fail(LDPC_B200_EINVAL, "quant_bits must be 0 (= 4) or 1..6");
    if (c.interleave_mod_type < 1 || LDPC_B200_N % c.interleave_mod_type) return fail(LDPC_B200_EINVAL, "InterleaveModType must divide N");
    if (c.puncture_tail < 0 || c.puncture_tail > LDPC_B200_N) return fail(LDPC_B200_EINVAL, "puncture_tail out of range");
    if (c.bf_mode < 0 || c.bf_mode > 3 || c.bf_max_iter < 0) return fail(LDPC_B200_EINVAL, "bad bf_mode / bf_max_iter");
    if (c.ef_elimination < 0 || c.ef_elimination > 2) return fail(LDPC_B200_EINVAL, "EF_ELIMINATION must be 0, 1 or 2");
    if (c.ef_elimination == 2 && c.decode_method == 2 && c.regular_col_weight != 3)
        return fail(LDPC_B200_EINVAL, "EF_ELIMINATION 2 erases weight-3 variable nodes: regular_col_weight must be 3");
    if (c.ef_floor_err_count < 0 || c.ef_floor_err_count > 127) return fail(LDPC_B200_EINVAL, "ef_floor_err_count must be in [0,127]");
    if (c.oms_floor_err_count < 0 || c.oms_floor_err_count > 255) return fail(LDPC_B200_EINVAL, "oms_floor_err_count must be in [0,255]");
    if (c.hard2_threshold < 1 || c.hard2_threshold > 31) return fail(LDPC_B200_EINVAL, "hard2_threshold must be in [1,31]");
    if (c.regular_col_weight < 1 || c.regular_col_weight > 12) return fail(LDPC_B200_EINVAL, "regular_col_weight out of range");
    if (c.dtbf_alpha < 0 || c.dtbf_alpha > 8 || c.dtbf_delta < 0) return fail(LDPC_B200_EINVAL, "dtbf_alpha / dtbf_delta out of range");
    const int m = c.decode_method;
    if ((m == 1 || m == 3 || m == 4) && ((int8_t)c.factor_1 < 0 || (int8_t)c.factor_2 < 0))
        return fail(LDPC_B200_EINVAL, "OMS Factor_1/Factor_2 must be non-negative");
    for (int it = 0; it < 6; ++it)
        for (int w = 0; w < 4; ++w)
            for (int a = 0; a < 8; ++a)
                if (c.v2c_lut[it][w][a] < 0 || c.v2c_lut[it][w][a] > 7 || c.v2c_lut_ef[it][w][a] < 0 || c.v2c_lut_ef[it][w][a] > 7)
                    return fail(LDPC_B200_EINVAL, "V2C LUT entries must be in [0,7] (4-bit messages)");
    return LDPC_B200_OK;
}


// Largest dynamic shared memory finalize_kernel can ask for (2B1C: hard + unsat + diff + hard2 per frame).  The attribute is
// per function and per DEVICE, not per handle: it is raised to this maximum once per device, so handles of different
// DecodeMethods coexist on one GPU whatever the order they were created in.
constexpr size_t kFinSmemMax = (size_t)32 * std::max(std::max(fin_layout_words(1, true), fin_layout_words(2, true)), fin_layout_words(3, true)) * sizeof(uint32_t);
int ensure_finalize_attr(int device) {
    static std::atomic<bool> done[64];
    if (device >= 0 && device < 64 && done[device].load(std::memory_order_acquire)) return LDPC_B200_OK;
    CUDA_TRY(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFinSmemMax));
    if (device >= 0 && device < 64) done[device].store(true, std::memory_order_release);
    return LDPC_B200_OK;
}

// NUMA node of a CUDA device and the CPUs of that node this process may run on (0 = unknown, or LDPC_B200_NUMA=0).
// Pinned buffers are allocated, and the staging threads run, on the socket the GPU's PCIe root is attached to: a DMA that
// crosses the inter-socket link runs at a fraction of the local rate (this is what kept the 2- and 4-GPU host-buffer figures
// of round 1 below the box's copy ceiling).
int device_numa_cpus(int device, cpu_set_t* set, int* node_out) {
    CPU_ZERO(set);
    if (node_out) *node_out = -1;
    const char* e = getenv("LDPC_B200_NUMA");
    if (e && atoi(e) == 0) return 0;
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (char* c = bdf; *c; ++c) *c = (char)tolower(*c);
    const int node = numa_node_of_pci(bdf);
    if (node_out) *node_out = node;
    return numa_node_cpus(node, set, sizeof(cpu_set_t));
}

void comm_destroy(FrameState& fs);  // frame_api.inl (NCCL is resolved with dlopen there)

void free_slot(Slot& s) {
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.final_hard) cudaFree(s.final_hard);
    if (s.snap) cudaFree(s.snap);
    if (s.grp_cnt) cudaFree(s.grp_cnt);
    if (s.first_zero) cudaFree(s.first_zero);
    if (s.work_counter) cudaFree(s.work_counter);
    if (s.d_bf) cudaFree(s.d_bf);
    if (s.d_its) cudaFree(s.d_its);
    if (s.d_conv) cudaFree(s.d_conv);
    if (s.h_bf) cudaFreeHost(s.h_bf);
    if (s.h_its) cudaFreeHost(s.h_its);
    if (s.h_conv) cudaFreeHost(s.h_conv);
    if (s.h_in_packed) cudaFreeHost(s.h_in_packed);
    if (s.h_out_packed) cudaFreeHost(s.h_out_packed);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_mid) cudaEventDestroy(s.ev_mid);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

// Streams, events and scratch of one chunk in flight, sized for h->chunk_groups groups.
int alloc_slot(ldpc_b200_handle* h, Slot& s) {
    const int cg = h->chunk_groups;
    const int mi = std::max(1, h->max_iteration_alloc);
    const size_t frames = (size_t)cg * 32;
    bool ok = true;
    ok = ok && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&s.ev_k0) == cudaSuccess && cudaEventCreate(&s.ev_k1) == cudaSuccess && cudaEventCreate(&s.ev_mid) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&s.d_in, frames * kN) == cudaSuccess;
    ok = ok && cudaMalloc(&s.d_out, frames * kN) == cudaSuccess;
    ok = ok && cudaMalloc(&s.final_hard, frames * h->planes * kHW * 4) == cudaSuccess;
    if (h->has_syndrome) {
        ok = ok && cudaMalloc(&s.snap, frames * mi * h->planes * kHW * 4) == cudaSuccess;
        ok = ok && cudaMalloc(&s.grp_cnt, (size_t)cg * mi * 4) == cudaSuccess;
    }
    ok = ok && cudaMalloc(&s.first_zero, frames * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&s.work_counter, sizeof(unsigned int)) == cudaSuccess;
    ok = ok && cudaMalloc(&s.d_bf, cg * 4) == cudaSuccess && cudaMalloc(&s.d_its, cg * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&s.d_conv, frames * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&s.h_bf, cg * 4) == cudaSuccess && cudaMallocHost(&s.h_its, cg * 4) == cudaSuccess;
    ok = ok && cudaMallocHost(&s.h_conv, frames * 4) == cudaSuccess;
    if (ok) ok = cudaEventRecord(s.ev_done, s.stream) == cudaSuccess;
    if (!ok) {
        const std::string msg = std::string("allocating decoder workspace: ") + cudaGetErrorString(cudaGetLastError());
        free_slot(s);
        return fail(LDPC_B200_ENOMEM, msg);
    }
    return LDPC_B200_OK;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Page-lock a pageable caller array once (opt-in, see ldpc_b200_handle::registered).  Failure is not an error: the array is
// then simply used as pageable memory.
void maybe_register(ldpc_b200_handle* h, const void* p, size_t bytes) {
    static const bool enabled = [] { const char* e = getenv("LDPC_B200_HOST_REGISTER"); return e && atoi(e) != 0; }();
    if (!enabled || !p || bytes == 0) return;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type != cudaMemoryTypeUnregistered) return;  // pinned / device already
    cudaGetLastError();
    for (auto it = h->registered.begin(); it != h->registered.end(); ++it)
        if (it->ptr == p) {  // same address, other size (or a registration that was lost): start over
            cudaHostUnregister(const_cast<void*>(p));
            cudaGetLastError();
            h->registered.erase(it);
            break;
        }
    if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) h->registered.push_back({p, bytes});
    else cudaGetLastError();
}

bool is_pinned_host_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Decode one chunk whose input is already on the device.  d_in: reference layout (packed_in = false) or native
// nibble layout; outputs to d_dec (reference layout bytes) and/or d_packed.
// gen != nullptr: fused producer -- the kernel synthesises the frames itself (d_in is ignored).
// want_info: the per-group / per-frame outputs of finalize_kernel (BF iterations, executed iterations, convergence
// iteration) are needed.  Without them DecodeMethod 0 writes its output straight from the decode kernel.
int run_chunk(ldpc_b200_handle* h, Slot& s, const void* d_in, bool packed_in, int8_t* d_dec, uint32_t* d_packed, int groups,
              const GenCore* gen = nullptr, bool want_info = true) {
    const ldpc_b200_config& c = h->cfg;
    const int frames = groups * 32;
    DecParams P;
    const bool mono = fill_dec_params(c, h->kind, h->planes, P);
    P.llr = packed_in ? nullptr : (const int8_t*)d_in;
    P.llr_packed = packed_in ? (const uint8_t*)d_in : nullptr;
    if (gen) {
        P.gen = *gen;
        P.gen_enable = 1;
    }
    const bool direct = h->kind == KIND_NMS && !want_info && getenv("LDPC_B200_NO_DIRECT_OUTPUT") == nullptr;
    if (direct) {
        P.direct_bytes = d_dec;
        P.direct_packed = d_dec ? nullptr : d_packed;
    }
    P.final_hard = s.final_hard;
    P.snap = s.snap;
    P.grp_cnt = s.grp_cnt;
    P.first_zero = s.first_zero;
    P.n_frames = frames;
    P.dbg = h->d_dbg;
    if (getenv("LDPC_B200_NO_SKEW")) P.no_skew = 1;
    if (getenv("LDPC_B200_EXP_NOLOAD")) P.exp_noload = 1;
    if (getenv("LDPC_B200_DEBUG_FAULT")) P.exp_fault = 1;
    P.work_counter = s.work_counter;
#if LDPC_PERSISTENT
    CUDA_TRY(cudaMemsetAsync(s.work_counter, 0, sizeof(unsigned int), s.stream));
#endif

    if (h->has_syndrome && c.max_iteration > 0)
        CUDA_TRY(cudaMemsetAsync(s.grp_cnt, 0, (size_t)groups * c.max_iteration * sizeof(uint32_t), s.stream));
    CUDA_TRY(cudaEventRecord(s.ev_k0, s.stream));
    CUDA_TRY(launch_decode_any(h->kind, mono, P, frames / 2, c.device, s.stream));
    CUDA_TRY(cudaEventRecord(s.ev_mid, s.stream));

    if (!direct) {
        FinParams F;
        memset(&F, 0, sizeof F);
        F.final_hard = s.final_hard;
        F.snap = s.snap;
        F.grp_cnt = s.grp_cnt;
        F.first_zero = s.first_zero;
        F.n_groups = groups;
        F.max_iter = c.max_iteration;
        F.planes = h->planes;
        F.has_syndrome = h->has_syndrome && c.max_iteration > 0;
        const int m = method_of(c);
        F.bf_mode = (m == 0 || m == 1) ? BF_NONE : c.bf_mode;
        F.bf_max_iter = c.bf_max_iter;
        F.L0 = c.dtbf_L0; F.L1 = c.dtbf_L1; F.delta = c.dtbf_delta; F.alpha = c.dtbf_alpha; F.rcw = c.regular_col_weight;
        // unrolled BF stage: weight-3 "regular" columns (the only weight-3 class of this code) and alpha in {0,1}
        F.dbg = h->d_dbg;
        F.wpf = h->fin_wpf;
        F.unsat_bufs = h->fin_unsat_bufs;
        F.fast_bf = c.regular_col_weight == 3 && c.dtbf_alpha <= 1 && c.dtbf_delta <= 8 && getenv("LDPC_B200_NO_FAST_BF") == nullptr;
        F.decoded = d_dec;
        F.hard_packed = d_packed;
        F.bf_iters = s.d_bf;
        F.its_per_group = s.d_its;
        F.conv_iter = s.d_conv;
        finalize_kernel<<<groups, kFinThreads, h->fin_smem, s.stream>>>(F);
        CUDA_TRY(cudaGetLastError());
        h->last_launches += 1;
    }
    CUDA_TRY(cudaEventRecord(s.ev_k1, s.stream));
    s.timing_pending = true;
    h->last_launches += 1;
    return LDPC_B200_OK;
}

int collect_timing(ldpc_b200_handle* h, Slot& s) {
    if (!s.timing_pending) return LDPC_B200_OK;
    CUDA_TRY(cudaEventSynchronize(s.ev_k1));
    float ms = 0.f, ms_dec = 0.f, ms_fin = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    CUDA_TRY(cudaEventElapsedTime(&ms_dec, s.ev_k0, s.ev_mid));
    CUDA_TRY(cudaEventElapsedTime(&ms_fin, s.ev_mid, s.ev_k1));
    h->last_kernel_ms += ms;
    h->last_decode_ms += ms_dec;
    h->last_finalize_ms += ms_fin;
    s.timing_pending = false;
    return LDPC_B200_OK;
}

// Host-side completion of the slot's drained chunk: the per-group outputs leave their pinned mirrors, the packed decisions
// are expanded into the caller's byte-per-bit array.  Runs when the slot is reused or the call drains, so that chunks keep
// overlapping even when the caller asks for the BF iteration counts (the `int` the reference's Decode_*() return).
void finish_chunk_info(Slot& s) {
    if (s.bf_dst) memcpy(s.bf_dst, s.h_bf, s.info_groups * sizeof(int32_t));
    if (s.its_dst) memcpy(s.its_dst, s.h_its, s.info_groups * sizeof(int32_t));
    if (s.conv_dst) memcpy(s.conv_dst, s.h_conv, (size_t)s.info_groups * 32 * sizeof(int32_t));
    s.bf_dst = s.its_dst = s.conv_dst = nullptr;
}
void finish_chunk(ldpc_b200_handle* h, Slot& s) {
    finish_chunk_info(s);
    if (!s.unpack_dst) return;
    host_unpack_bits(h->pool, s.h_out_packed, s.unpack_dst, s.unpack_frames);
    s.unpack_dst = nullptr;
    s.unpack_frames = 0;
}

int decode_impl_inner(ldpc_b200_handle* h, const void* in, bool packed_in, int8_t* dec, uint32_t* packed_out, int n_groups,
                      int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter);

int decode_impl(ldpc_b200_handle* h, const void* in, bool packed_in, int8_t* dec, uint32_t* packed_out, int n_groups,
                int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    const int rc = decode_impl_inner(h, in, packed_in, dec, packed_out, n_groups, bf_iters, its_per_group, conv_iter);
    if (rc != LDPC_B200_OK) {
        // A call that fails half-way must not leave copies in flight into the caller's arrays after it has returned:
        // drain every slot and drop the pending host-side completions (the error text of the failure is kept).
        const std::string msg = g_last_error;
        for (auto* v : {&h->slots, &h->dslots})
            for (auto& s : *v) {
                if (s.stream) cudaStreamSynchronize(s.stream);
                s.unpack_dst = nullptr;
                s.bf_dst = s.its_dst = s.conv_dst = nullptr;
                s.timing_pending = false;
            }
        cudaGetLastError();
        g_last_error = msg;
    }
    return rc;
}

int decode_impl_inner(ldpc_b200_handle* h, const void* in, bool packed_in, int8_t* dec, uint32_t* packed_out, int n_groups,
                      int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter) {
    if (n_groups < 0 || !in || (!dec && !packed_out && n_groups > 0)) return fail(LDPC_B200_EINVAL, "bad decode arguments");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->last_kernel_ms = h->last_decode_ms = h->last_finalize_ms = 0.f;
    h->last_launches = 0;
    if (n_groups == 0) return LDPC_B200_OK;
    if ((uintptr_t)in & 3) return fail(LDPC_B200_EINVAL, "input pointer must be 4-byte aligned");
    if (dec && ((uintptr_t)dec & 15)) return fail(LDPC_B200_EINVAL, "decodedBits pointer must be 16-byte aligned");
    const bool in_dev = is_device_ptr(in);
    const bool out_dev = is_device_ptr(dec ? (const void*)dec : (const void*)packed_out);
    const size_t in_group_bytes = packed_in ? (size_t)32 * kN / 2 : (size_t)32 * kN;
    const size_t out_group_bytes = dec ? (size_t)32 * kN : (size_t)32 * kHW * 4;
    if (!in_dev) maybe_register(h, in, (size_t)n_groups * in_group_bytes);
    if (!out_dev) maybe_register(h, dec ? (const void*)dec : (const void*)packed_out, (size_t)n_groups * out_group_bytes);
    const int ns = (int)h->slots.size();
    const bool want_info = bf_iters || its_per_group || conv_iter;
    // Host staging (host_pack.h): byte-per-bit decisions cross PCIe as bits and are expanded into the caller's array by
    // the handle's host threads while the next chunks decode; optionally the int8 LLRs cross as nibbles.
    const bool stage_out = h->pool && h->stage_out && dec && !out_dev;
    const bool stage_in = h->pool && h->stage_in && !packed_in && !in_dev;
    // With the library's default chunking, calls with host arrays run in small chunks on the handle's slots (default 6) so that
    // copies, host staging and kernels of different chunks overlap; device-resident calls keep the large chunk.  Measured on
    // B200: staged 64 x 4 = 41.2, 128 x 5 = 41.0, 128 x 3 = 39.4, 32 x 6 = 38.1 Gbit/s (profiles/r02_e2e_chunks_exp8.log), with
    // one direct slot 128 x 4 = 43.0, 64 x 4 = 42.5 (r02_e2e_hybrid_sweep.log); hybrid with bit output 64 x 6 = 51.0 / 46.4,
    // 128 x 6 = 49.5 / 44.7, 128 x 4 = 47.2 on two boxes (r02_e2e_semidirect_exp11{,b}.log);
    // copied as they are 32 x 6 = 36.9, 64 x 4 = 32.4, 128 x 3 = 33.2 of a 40.2 Gbit/s copy ceiling (r02_e2e_exp2_modes.log).
    const int host_chunk = (stage_in || stage_out) ? 64 : 32;
    const int chunk = (h->chunk_default && (!in_dev || !out_dev)) ? std::min(h->chunk_groups, host_chunk) : h->chunk_groups;
    const size_t cap_frames = (size_t)chunk * 32;  // pinned staging mirrors are sized for the chunks this path uses
    h->last_h2d_bytes = h->last_d2h_bytes = 0;
    for (auto& s : h->slots) {  // a call that failed half-way must not leak its pending completions
        s.unpack_dst = nullptr;
        s.bf_dst = s.its_dst = s.conv_dst = nullptr;
    }
    // Hybrid host-buffer path (LDPC_B200_HYBRID = number of "direct" slots, default 1, 0 = off): whenever a direct slot is idle
    // the next chunk's LLRs are copied as they are by the copy engine, otherwise the host threads pack them to nibbles -- the two
    // routes use different resources (PCIe link vs host threads / host DRAM) and the split adapts by itself.  The decisions of
    // EVERY chunk come back as bits and are expanded by the host threads: expanding costs the host 22 KB of DRAM traffic per
    // frame against 35 KB for packing, and a byte-per-bit copy would load the link's other direction for nothing.  Needs a pinned
    // fixInput array.  Measured on two 16-core B200 boxes (profiles/r02_e2e_semidirect_exp11{,b}.log): 49.8-51.0 and 46.4-46.6
    // Gbit/s, against 43.4-44.4 and 44.1-44.8 when the direct slot also copies its decisions as bytes (LDPC_B200_HYBRID_OUT_BITS=0,
    // the first form of this path, which needs decodedBits pinned too) and 42.8-44.8 / 43.0-43.5 all-staged.  More than one direct
    // slot is slower (45-47): the big copies then delay the staged chunks' small ones and the host threads wait for their slots.
    const char* e_hyb = getenv("LDPC_B200_HYBRID");
    const int n_direct = e_hyb ? std::max(0, std::min(6, atoi(e_hyb))) : 1;
    const char* e_hout = getenv("LDPC_B200_HYBRID_OUT_BITS");
    const bool direct_out_bits = !(e_hout && atoi(e_hout) == 0);
    const bool hybrid = n_direct > 0 && stage_in && stage_out && !in_dev && !out_dev && n_groups >= 4 * chunk &&
                        is_pinned_host_ptr(in) && (direct_out_bits || is_pinned_host_ptr(dec));
    // LDPC_B200_HYBRID_CHUNK: groups per direct chunk when it should differ from the staged chunk (no gain measured)
    const char* e_hchunk = getenv("LDPC_B200_HYBRID_CHUNK");
    const int dchunk = e_hchunk ? std::max(1, std::min(chunk, atoi(e_hchunk))) : chunk;
    if (hybrid && h->dslots.empty()) {
        h->dslots.resize(n_direct);
        for (auto& d : h->dslots) {
            const int rc = alloc_slot(h, d);
            if (rc) {
                for (auto& t : h->dslots) free_slot(t);
                h->dslots.clear();
                return rc;
            }
        }
    }
    for (auto& d : h->dslots) {
        d.unpack_dst = nullptr;
        d.bf_dst = d.its_dst = d.conv_dst = nullptr;
    }
    h->last_direct_chunks = h->last_staged_chunks = 0;
    int chunk_idx = 0;
    for (int g0 = 0, groups = 0; g0 < n_groups; g0 += groups) {
        Slot* sp = nullptr;
        bool direct = false;
        if (hybrid)
            for (auto& d : h->dslots) {
                const cudaError_t q = cudaEventQuery(d.ev_done);
                if (q == cudaSuccess) { sp = &d; direct = true; break; }
                if (q != cudaErrorNotReady) return fail(LDPC_B200_ECUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
                cudaGetLastError();
            }
        groups = std::min(direct ? dchunk : chunk, n_groups - g0);
        if (!sp) {
            sp = &h->slots[chunk_idx++ % ns];
            // the slot's previous chunk must have fully drained (its staging buffers are about to be reused)
            CUDA_TRY(cudaEventSynchronize(sp->ev_done));
        }
        Slot& s = *sp;
        // host side of the slot's previous chunk: small outputs now; its bit expansion is fused with this chunk's nibble
        // packing below (one pass over the host threads) when both exist
        finish_chunk_info(s);
        const bool fuse_stage = stage_in && !direct && s.unpack_dst != nullptr;
        if (!fuse_stage) finish_chunk(h, s);
        int rc = collect_timing(h, s);
        if (rc) return rc;
        (direct ? h->last_direct_chunks : h->last_staged_chunks) += 1;
        const uint8_t* src = (const uint8_t*)in + (size_t)g0 * in_group_bytes;
        const void* d_in = src;
        bool chunk_packed = packed_in;
        if (!in_dev) {
            bool nibbles = false;
            if (stage_in && !direct) {
                if (!s.h_in_packed) {
                    ScopedAffinity local(h->numa_ncpu ? &h->numa_cpus : nullptr, sizeof(cpu_set_t));
                    CUDA_TRY(cudaMallocHost(&s.h_in_packed, cap_frames * (kN / 2)));
                }
                // false: a value outside [-8,7] (6-bit quantiser range) -- the chunk then goes up as bytes
                if (fuse_stage) {
                    nibbles = host_stage_both(h->pool, (const int8_t*)src, s.h_in_packed, groups, s.h_out_packed, s.unpack_dst, s.unpack_frames);
                    s.unpack_dst = nullptr;
                    s.unpack_frames = 0;
                } else {
                    nibbles = host_pack_llr(h->pool, (const int8_t*)src, s.h_in_packed, groups);
                }
            }
            if (nibbles) {
                CUDA_TRY(cudaMemcpyAsync(s.d_in, s.h_in_packed, (size_t)groups * 32 * (kN / 2), cudaMemcpyHostToDevice, s.stream));
                h->last_h2d_bytes += (uint64_t)groups * 32 * (kN / 2);
                chunk_packed = true;
            } else {
                CUDA_TRY(cudaMemcpyAsync(s.d_in, src, (size_t)groups * in_group_bytes, cudaMemcpyHostToDevice, s.stream));
                h->last_h2d_bytes += (uint64_t)groups * in_group_bytes;
            }
            d_in = s.d_in;
        }
        uint8_t* dst = (dec ? (uint8_t*)dec : (uint8_t*)packed_out) + (size_t)g0 * out_group_bytes;
        if (stage_out && (!direct || direct_out_bits)) {
            if (!s.h_out_packed) {
                ScopedAffinity local(h->numa_ncpu ? &h->numa_cpus : nullptr, sizeof(cpu_set_t));
                CUDA_TRY(cudaMallocHost(&s.h_out_packed, cap_frames * kHW * sizeof(uint32_t)));
            }
            rc = run_chunk(h, s, d_in, chunk_packed, nullptr, (uint32_t*)s.d_out, groups, nullptr, want_info);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(s.h_out_packed, s.d_out, (size_t)groups * 32 * kHW * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
            h->last_d2h_bytes += (uint64_t)groups * 32 * kHW * sizeof(uint32_t);
            s.unpack_dst = (int8_t*)dst;
            s.unpack_frames = groups * 32;
        } else {
            void* d_out = out_dev ? (void*)dst : (void*)s.d_out;
            rc = run_chunk(h, s, d_in, chunk_packed, dec ? (int8_t*)d_out : nullptr, dec ? nullptr : (uint32_t*)d_out, groups, nullptr, want_info);
            if (rc) return rc;
            if (!out_dev) {
                CUDA_TRY(cudaMemcpyAsync(dst, s.d_out, (size_t)groups * out_group_bytes, cudaMemcpyDeviceToHost, s.stream));
                h->last_d2h_bytes += (uint64_t)groups * out_group_bytes;
            }
        }
        // small per-group outputs: through pinned mirrors, copied out after the stream drains
        if (bf_iters) CUDA_TRY(cudaMemcpyAsync(s.h_bf, s.d_bf, groups * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        if (its_per_group) CUDA_TRY(cudaMemcpyAsync(s.h_its, s.d_its, groups * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        if (conv_iter) CUDA_TRY(cudaMemcpyAsync(s.h_conv, s.d_conv, (size_t)groups * 32 * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(cudaEventRecord(s.ev_done, s.stream));
        s.bf_dst = bf_iters ? bf_iters + g0 : nullptr;
        s.its_dst = its_per_group ? its_per_group + g0 : nullptr;
        s.conv_dst = conv_iter ? conv_iter + (size_t)g0 * 32 : nullptr;
        s.info_groups = groups;
    }
    for (auto& d : h->dslots) {
        CUDA_TRY(cudaStreamSynchronize(d.stream));
        finish_chunk(h, d);
        int rc = collect_timing(h, d);
        if (rc) return rc;
    }
    // drain, oldest chunk first
    for (int k = 0; k < ns; ++k) {
        Slot& s = h->slots[(chunk_idx + k) % ns];
        CUDA_TRY(cudaStreamSynchronize(s.stream));
        finish_chunk(h, s);
        int rc = collect_timing(h, s);
        if (rc) return rc;
    }
    return LDPC_B200_OK;
}

}  // namespace

extern "C" {

const char* ldpc_b200_version(void) { return "ldpc_b200 0.1 (sm_100a)"; }
const char* ldpc_b200_last_error(void) { return g_last_error.c_str(); }

int ldpc_b200_default_config(ldpc_b200_config* c, int method, int lut_variant) {
    if (!c) return fail(LDPC_B200_EINVAL, "null config");
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->abi_version = LDPC_B200_ABI_VERSION;
    // Profile.txt as shipped
    c->snr_start = 3.0f; c->snr_pass = 0.1f; c->snr_end = 5.0f;
    c->decode_method = method;
    c->max_iteration = 6;
    c->mod_type = 2;
    c->interleave_mod_type = 1;
    c->factor_1 = 1; c->factor_2 = 6;
    if (method == LDPC_B200_NMS) { c->factor_1 = 26; c->factor_2 = 26; }  // README.md:22
    c->nb_frames = 32;
    c->scale = (method == LDPC_B200_FAID_2B1C) ? 12.5f : 13.0f;           // README.md:22
    c->Z = 256;
    c->puncture_tail = 384;    // CLDPC.cpp:270-272
    c->code_rate = 0.8444444;  // CLDPC.cpp:4780
    method_constants(c, (method < 0 || method > 5) ? 0 : method, lut_variant);
    c->device = 0;
    c->n_streams = 6;
    c->chunk_groups = 0;
    return LDPC_B200_OK;
}

int ldpc_b200_read_profile(const char* path, ldpc_b200_config* c, int lut_variant) {
    if (!path || !c) return fail(LDPC_B200_EINVAL, "null argument");
    std::ifstream fin(path);
    if (!fin.is_open()) return fail(LDPC_B200_EIO, std::string("Cannot open Profile: ") + path);
    // token order of ReadProfile, CTool.cpp:597-616
    std::string rub, fname;
    float snr_start, snr_pass, snr_end, scale;
    int method, max_it, mod, il, f1, f2, nfr, Z;
    fin >> rub >> rub;
    fin >> rub >> snr_start;
    fin >> rub >> snr_pass;
    fin >> rub >> snr_end;
    fin >> rub >> method;
    fin >> rub >> max_it;
    fin >> rub >> rub;
    fin >> rub >> mod;
    fin >> rub >> il;
    fin >> rub >> rub;
    fin >> rub >> f1;
    fin >> rub >> f2;
    fin >> rub >> nfr;
    fin >> rub >> scale;
    fin >> rub >> rub;
    fin >> rub >> fname;
    fin >> rub >> Z;
    if (fin.fail()) return fail(LDPC_B200_EIO, std::string("Malformed Profile: ") + path);
    if (c->struct_size != sizeof *c) {
        int rc = ldpc_b200_default_config(c, method, lut_variant);
        if (rc) return rc;
    } else if (c->decode_method != method) {
        method_constants(c, (method < 0 || method > 5) ? 0 : method, lut_variant);
    }
    c->snr_start = snr_start; c->snr_pass = snr_pass; c->snr_end = snr_end;
    c->decode_method = method; c->max_iteration = max_it;
    c->mod_type = mod; c->interleave_mod_type = il;
    c->factor_1 = f1; c->factor_2 = f2;
    c->nb_frames = nfr; c->scale = scale; c->Z = Z;
    return LDPC_B200_OK;
}

int ldpc_b200_create(const ldpc_b200_config* cfg, ldpc_b200_handle** out) {
    if (!cfg || !out) return fail(LDPC_B200_EINVAL, "null argument");
    *out = nullptr;
    int rc = validate(*cfg);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(LDPC_B200_ENODEV, "no CUDA device available; this engine has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(LDPC_B200_EINVAL, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(LDPC_B200_ENODEV, "device is not sm_100 class (kernels are built for sm_100a only)");

    ldpc_b200_handle* h = new ldpc_b200_handle();
    h->cfg = *cfg;
    const int m = method_of(*cfg);
    h->kind = kind_of(*cfg, getenv("LDPC_B200_NO_FAID_FAST") == nullptr);  // env switch: A/B and tests of the general FAID path
    h->has_syndrome = m != 0;
    const int bf_mode = (m == 0 || m == 1) ? BF_NONE : cfg->bf_mode;
    h->planes = (bf_mode == BF_2B1C) ? 2 : 1;
    rc = upload_tables(*cfg);
    if (rc) { delete h; return rc; }

    // finalize kernel shared memory: per frame hard [+ unsat + diff [+ hard2]]
    const bool do_bf = bf_mode != BF_NONE && cfg->bf_max_iter > 0;
    if (cudaMalloc(&h->d_dbg, 2 * sizeof(unsigned long long)) != cudaSuccess || cudaMemset(h->d_dbg, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) {
        delete h;
        return fail(LDPC_B200_ENOMEM, "allocating the bounds-check record");
    }
    h->fin_wpf = fin_layout_words(bf_mode, do_bf);
    h->fin_unsat_bufs = fin_unsat_bufs(bf_mode, do_bf);
    h->fin_smem = (size_t)32 * h->fin_wpf * sizeof(uint32_t);
    h->max_iteration_alloc = cfg->max_iteration;
    rc = ensure_finalize_attr(cfg->device);
    if (rc) { delete h; return rc; }

    // chunking: bound the scratch (snapshots dominate) to ~2 GiB per slot
    const int mi = std::max(1, cfg->max_iteration);
    const size_t per_group = (size_t)32 * kN * 2 + (size_t)32 * h->planes * kHW * 4 * (1 + (h->has_syndrome ? mi : 0));
    int cg = cfg->chunk_groups > 0 ? cfg->chunk_groups : 1024;
    const size_t budget = (size_t)2 << 30;
    cg = (int)std::max<size_t>(1, std::min<size_t>((size_t)cg, budget / per_group));
    h->chunk_groups = cg;
    h->chunk_default = cfg->chunk_groups <= 0;
    const int ns = std::max(1, std::min(8, cfg->n_streams));
    h->slots.resize(ns);
    for (auto& s : h->slots) {
        rc = alloc_slot(h, s);
        if (rc) {
            const std::string msg = g_last_error;
            for (auto& t : h->slots) free_slot(t);
            delete h;
            return fail(rc, msg);
        }
    }
    // Host staging (host_pack.h).  Threads: this rank's share of the host's cores (hardware threads / LOCAL_WORLD_SIZE, at most
    // the CPUs of the GPU's NUMA node, at most 64), pinned to that node.  Defaults: decisions return as bits and are expanded by
    // the threads whenever the rank has >= 4 of them; LLRs are nibble-packed by the threads only when this process has the host
    // to itself (>= 8 cores) -- packing costs the host's memory system 35 KB per frame against 18 KB for a copy as it is, and
    // with several links busy the box is bound by exactly that (2 GPUs, 12 threads each, same call: 59.0-59.3 Gbit/s with bits
    // out, 50.0-51.5 with both arrays copied as they are, 51.3-54.3 with the single-rank hybrid; profiles/r02_bench_2gpu_exp13_*.json).
    // Overrides: LDPC_B200_HOST_THREADS (0 = off), LDPC_B200_STAGE_OUT, LDPC_B200_STAGE_IN, LDPC_B200_NUMA (0 = no placement).
    {
        h->numa_ncpu = device_numa_cpus(cfg->device, &h->numa_cpus, &h->numa_node);
        const char* e_thr = getenv("LDPC_B200_HOST_THREADS");
        const char* e_lws = getenv("LOCAL_WORLD_SIZE");
        const char* e_out = getenv("LDPC_B200_STAGE_OUT");
        const char* e_in = getenv("LDPC_B200_STAGE_IN");
        const int ranks = e_lws ? std::max(1, atoi(e_lws)) : 1;
        int cores = std::max(1, (int)std::thread::hardware_concurrency() / ranks);
        if (h->numa_ncpu > 0) cores = std::min(cores, h->numa_ncpu);
        const int n_thr = e_thr ? atoi(e_thr) : std::min(64, cores);
        h->stage_out = e_out ? atoi(e_out) != 0 : cores >= 4;
        h->stage_in = e_in ? atoi(e_in) != 0 : (ranks == 1 && cores >= 8);
        if (n_thr > 0 && (h->stage_out || h->stage_in))
            h->pool = host_pool_create(n_thr, h->numa_ncpu ? &h->numa_cpus : nullptr, sizeof(cpu_set_t));
    }
    rc = frame_state_init(h->fs, *cfg);
    if (rc) {
        frame_state_free(h->fs);
        host_pool_destroy(h->pool);
        for (auto& t : h->slots) free_slot(t);
        delete h;
        return fail(rc, "allocating frame-generation workspace");
    }
    *out = h;
    return LDPC_B200_OK;
}

int ldpc_b200_destroy(ldpc_b200_handle* h) {
    if (!h) return LDPC_B200_OK;
    cudaSetDevice(h->cfg.device);
    for (auto* v : {&h->slots, &h->dslots})
        for (auto& s : *v) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            free_slot(s);
        }
    for (auto& r : h->registered) cudaHostUnregister(const_cast<void*>(r.ptr));
    cudaGetLastError();
    comm_destroy(h->fs);
    frame_state_free(h->fs);
    host_pool_destroy(h->pool);
    if (h->d_dbg) cudaFree(h->d_dbg);
    delete h;
    return LDPC_B200_OK;
}

int ldpc_b200_set_factors(ldpc_b200_handle* h, int f1, int f2) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    ldpc_b200_config c = h->cfg;
    c.factor_1 = f1;
    c.factor_2 = f2;
    int rc = validate(c);
    if (rc) return rc;
    h->cfg = c;
    return LDPC_B200_OK;
}

int ldpc_b200_set_max_iteration(ldpc_b200_handle* h, int mi) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (mi < 0 || mi > h->max_iteration_alloc)
        return fail(LDPC_B200_EINVAL, "max_iteration must be in [0, the value the handle was created with]: the scratch is sized for that");
    h->cfg.max_iteration = mi;
    return LDPC_B200_OK;
}

int ldpc_b200_decode(ldpc_b200_handle* h, const int8_t* fixInput, int8_t* decodedBits, int n_groups, int32_t* bf_iters,
                     int32_t* its_per_group, int32_t* conv_iter) {
    return decode_impl(h, fixInput, false, decodedBits, nullptr, n_groups, bf_iters, its_per_group, conv_iter);
}

int ldpc_b200_decode_packed(ldpc_b200_handle* h, const uint8_t* llr_packed, uint32_t* hard_packed, int n_groups,
                            int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter) {
    return decode_impl(h, llr_packed, true, nullptr, hard_packed, n_groups, bf_iters, its_per_group, conv_iter);
}

int ldpc_b200_last_timing(ldpc_b200_handle* h, float* kernel_ms, int32_t* launches) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (kernel_ms) *kernel_ms = h->last_kernel_ms;
    if (launches) *launches = h->last_launches;
    return LDPC_B200_OK;
}

int ldpc_b200_last_timing_detail(ldpc_b200_handle* h, float* decode_ms, float* finalize_ms) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (decode_ms) *decode_ms = h->last_decode_ms;
    if (finalize_ms) *finalize_ms = h->last_finalize_ms;
    return LDPC_B200_OK;
}

int ldpc_b200_host_staging(ldpc_b200_handle* h, int32_t* threads, int32_t* stage_in, int32_t* stage_out, uint64_t* last_h2d_bytes,
                           uint64_t* last_d2h_bytes) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (threads) *threads = host_pool_threads(h->pool);
    if (stage_in) *stage_in = h->pool && h->stage_in;
    if (stage_out) *stage_out = h->pool && h->stage_out;
    if (last_h2d_bytes) *last_h2d_bytes = h->last_h2d_bytes;
    if (last_d2h_bytes) *last_d2h_bytes = h->last_d2h_bytes;
    return LDPC_B200_OK;
}

int ldpc_b200_debug_bounds(ldpc_b200_handle* h, int32_t* compiled_in, uint64_t* violations, uint64_t* first) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long rec[2] = {0, 0};
    CUDA_TRY(cudaMemcpy(rec, h->d_dbg, sizeof rec, cudaMemcpyDeviceToHost));
    if (compiled_in) *compiled_in = LDPC_DEBUG_BOUNDS;
    if (violations) *violations = rec[0];
    if (first) *first = rec[1];
    return LDPC_B200_OK;
}

int ldpc_b200_last_routing(ldpc_b200_handle* h, int32_t* staged_chunks, int32_t* direct_chunks) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (staged_chunks) *staged_chunks = h->last_staged_chunks;
    if (direct_chunks) *direct_chunks = h->last_direct_chunks;
    return LDPC_B200_OK;
}

int ldpc_b200_host_placement(ldpc_b200_handle* h, int32_t* numa_node, int32_t* numa_cpus) {
    if (!h) return fail(LDPC_B200_EINVAL, "null handle");
    if (numa_node) *numa_node = h->numa_node;
    if (numa_cpus) *numa_cpus = h->numa_ncpu;
    return LDPC_B200_OK;
}

int ldpc_b200_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return fail(LDPC_B200_EINVAL, "null argument");
    // pages are placed on the NUMA node of the CURRENT CUDA device (cudaSetDevice before allocating; see device_numa_cpus)
    int dev = 0;
    cpu_set_t cpus;
    int ncpu = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) ncpu = device_numa_cpus(dev, &cpus, nullptr);
    else cudaGetLastError();
    ScopedAffinity local(ncpu ? &cpus : nullptr, sizeof(cpu_set_t));
    CUDA_TRY(cudaMallocHost(ptr, bytes));
    return LDPC_B200_OK;
}
int ldpc_b200_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return LDPC_B200_OK;
}

}  // extern "C"

// frame generation / encoder / counters / NCCL entry points
#include "frame_api.inl"
