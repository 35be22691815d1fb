// decode_launch.h -- launchers of decode_pair_kernel<KIND, MONO>, one translation unit per kernel kind (decode_inst.cu is
// compiled once per KIND with -DLDPC_INST_KIND=k so that the seven ~100 KB fully unrolled kernels build in parallel).
#pragma once
#include <cuda_runtime.h>

#include "decode_kernels.cuh"

namespace ldpc {

// Sets the function attributes (once per device, thread-safe) and launches.  Returns the CUDA error of either.
#define LDPC_DECL_LAUNCH(K) cudaError_t launch_decode_kind##K(bool mono, const DecParams& P, int n_pairs, int device, cudaStream_t st);
LDPC_DECL_LAUNCH(0) LDPC_DECL_LAUNCH(1) LDPC_DECL_LAUNCH(2) LDPC_DECL_LAUNCH(3) LDPC_DECL_LAUNCH(4) LDPC_DECL_LAUNCH(5) LDPC_DECL_LAUNCH(6)
#undef LDPC_DECL_LAUNCH

inline cudaError_t launch_decode_any(int kind, bool mono, const DecParams& P, int n_pairs, int device, cudaStream_t st) {
    switch (kind) {
    case KIND_NMS: return launch_decode_kind0(mono, P, n_pairs, device, st);
    case KIND_OMS: return launch_decode_kind1(mono, P, n_pairs, device, st);
    case KIND_FAID: return launch_decode_kind2(mono, P, n_pairs, device, st);
    case KIND_FAID_EF: return launch_decode_kind3(mono, P, n_pairs, device, st);
    case KIND_FAID_M: return launch_decode_kind4(mono, P, n_pairs, device, st);
    case KIND_FAID_EF_M: return launch_decode_kind5(mono, P, n_pairs, device, st);
    case KIND_FAID_ER: return launch_decode_kind6(mono, P, n_pairs, device, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace ldpc
