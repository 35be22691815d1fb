// decode_inst.cu -- instantiation + launcher of decode_pair_kernel for ONE kernel kind (-DLDPC_INST_KIND=k).
#ifndef LDPC_INST_KIND
#error "compile with -DLDPC_INST_KIND=<0..6>"
#endif
#include <algorithm>
#include <atomic>

#include "decode_launch.h"

namespace ldpc {
namespace {

template <int KIND, bool MONO>
cudaError_t launch_one(const DecParams& P, int n_pairs, int device, cudaStream_t st) {
    const size_t smem = decode_smem_bytes(KIND);  // APP words of the frame pairs + message words of the shared-memory-resident layers
    // function attributes are per device; a process may hold handles on several GPUs, driven from several host threads
    // (the reference's model: one object set per pthread, CSimulate.cpp:218-278)
    static std::atomic<bool> attr_set[64];
    if (device < 0 || device >= 64 || !attr_set[device].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(decode_pair_kernel<KIND, MONO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(decode_pair_kernel<KIND, MONO>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        if (device >= 0 && device < 64) attr_set[device].store(true, std::memory_order_release);
    }
    // persistent: one CTA per SM (1 CTA/SM is what the 227 KB of shared memory allow), never more CTAs than work
    static std::atomic<int> sm_count[64];
    int sms = (device >= 0 && device < 64) ? sm_count[device].load(std::memory_order_relaxed) : 0;
    if (sms <= 0) {
        cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e != cudaSuccess) return e;
        if (device >= 0 && device < 64) sm_count[device].store(sms, std::memory_order_relaxed);
    }
    const int grid = std::max(1, LDPC_PERSISTENT ? std::min((n_pairs + kPairsPerCta - 1) / kPairsPerCta, sms) : (n_pairs + kPairsPerCta - 1) / kPairsPerCta);
    decode_pair_kernel<KIND, MONO><<<grid, kThreads * kPairsPerCta, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

#define LDPC_CAT2(a, b) a##b
#define LDPC_CAT(a, b) LDPC_CAT2(a, b)
cudaError_t LDPC_CAT(launch_decode_kind, LDPC_INST_KIND)(bool mono, const DecParams& P, int n_pairs, int device, cudaStream_t st) {
    // MONO only matters for the min-sum kinds (single-instruction is-min select when cste_1 >= cste_2)
#if !LDPC_FP16_SELECT
    // integer select: MONO = "cste_1 >= cste_2 for every reachable pair of minima" (single-instruction is-min select)
    if (LDPC_INST_KIND <= KIND_OMS && !mono) return launch_one<LDPC_INST_KIND, false>(P, n_pairs, device, st);
#elif LDPC_INST_KIND == 0
    // fp16 select is valid for any cste order, so the NMS kernel uses MONO for "both factors in [0, 2114]": the scaling
    // (min * factor) >> 5 then cannot wrap 16 bits and runs on both halves at once (nms_scale16)
    if (!P.nms_fast) return launch_one<LDPC_INST_KIND, false>(P, n_pairs, device, st);
#endif
    (void)mono;
    return launch_one<LDPC_INST_KIND, true>(P, n_pairs, device, st);
}

}  // namespace ldpc
