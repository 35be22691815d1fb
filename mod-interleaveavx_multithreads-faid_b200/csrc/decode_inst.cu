// decode_inst.cu -- instantiation + launcher of decode_pair_kernel for ONE kernel kind (-DLDPC_INST_KIND=k).
#ifndef LDPC_INST_KIND
#error "compile with -DLDPC_INST_KIND=<0..6>"
#endif
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
    decode_pair_kernel<KIND, MONO><<<(n_pairs + kPairsPerCta - 1) / kPairsPerCta, kThreads * kPairsPerCta, smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

#define LDPC_CAT2(a, b) a##b
#define LDPC_CAT(a, b) LDPC_CAT2(a, b)
cudaError_t LDPC_CAT(launch_decode_kind, LDPC_INST_KIND)(bool mono, const DecParams& P, int n_pairs, int device, cudaStream_t st) {
    // MONO only matters for the min-sum kinds (single-instruction is-min select when cste_1 >= cste_2)
#if !LDPC_FP16_SELECT  // the fp16-pipe select is valid for any cste order: both values of MONO would be the same code
    if (LDPC_INST_KIND <= KIND_OMS && !mono) return launch_one<LDPC_INST_KIND, false>(P, n_pairs, device, st);
#endif
    (void)mono;
    return launch_one<LDPC_INST_KIND, true>(P, n_pairs, device, st);
}

}  // namespace ldpc
