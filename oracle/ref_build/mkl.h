/* Minimal stand-in for Intel MKL so the reference translation units compile with g++.
 * TEST INFRASTRUCTURE ONLY (oracle build).  The reference uses MKL for (a) the MKL_Complex8 POD type
 * (CChannel.h:31, CModulate.h:32) and (b) the MT2203 Gaussian stream of the BPSK-only channel
 * (CChannel.cpp:49,68,105).  (b) is out of scope: the stub returns zeros, so
 * BPSKAWGNChannel is NOT usable through this oracle build. */
#ifndef LDPC_ORACLE_MKL_STUB_H
#define LDPC_ORACLE_MKL_STUB_H
typedef struct { float real; float imag; } MKL_Complex8;
typedef void* VSLStreamStatePtr;
#define VSL_STATUS_OK 0
#define VSL_BRNG_MT2203 0
static inline int vslNewStream(VSLStreamStatePtr* s, int, int) { *s = 0; return VSL_STATUS_OK; }
static inline int vslDeleteStream(VSLStreamStatePtr*) { return VSL_STATUS_OK; }
static inline int vsRngGaussian(int, VSLStreamStatePtr, long n, float* r, float, float) {
    for (long i = 0; i < n; ++i) r[i] = 0.0f;
    return VSL_STATUS_OK;
}
#endif
