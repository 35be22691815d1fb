// Micro-benchmark: per-SM issue rate of the integer / packed / half2 instructions the
// LDPC decoder kernels can be built from, on sm_100a.  Test infrastructure, not product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
// Output: one JSON line per op: warp-instructions per clock per SM (4.0 = issue limit).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define ILP 8
#define INNER 64

template <class OP>
__global__ void __launch_bounds__(1024) bench(unsigned* out, unsigned seed, int iters, long long* cyc) {
    unsigned x[ILP];
    unsigned y = seed ^ (threadIdx.x * 2654435761u), z = (seed >> 3) + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = seed + i * 0x01010101u + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
            unsigned t[ILP];
            // second operand comes from a neighbouring chain so that no op is idempotent/foldable
#pragma unroll
            for (int i = 0; i < ILP; ++i) t[i] = OP::op(x[i], x[(i + 3) % ILP], z);
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = t[i];
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

#define DEFOP(name, n_instr, expr)                                                     \
    struct name {                                                                      \
        static constexpr int n = n_instr;                                              \
        static __device__ __forceinline__ unsigned op(unsigned x, unsigned y, unsigned z) { \
            expr;                                                                      \
        }                                                                              \
    };

static __device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }
static __device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2*>(&u); }

DEFOP(op_iadd3, 1, return x + y + z)
DEFOP(op_lop3, 1, return (x & y) ^ z)
DEFOP(op_imad, 1, return x * y + z)
DEFOP(op_shf, 1, return __funnelshift_l(x, y, 7))
DEFOP(op_prmt, 1, return __byte_perm(x, y, 0x9180))
DEFOP(op_vimnmx32, 1, return (unsigned)max((int)x, (int)y))
DEFOP(op_vimnmx3_32, 1, return (unsigned)__vimax3_s32((int)x, (int)y, (int)z))
DEFOP(op_viaddmnmx32, 1, return (unsigned)__viaddmax_s32((int)x, (int)y, (int)z))
DEFOP(op_vimnmx16x2, 1, return __vmaxs2(x, y))
DEFOP(op_vimnmx3_16x2, 1, return __vimax3_s16x2(x, y, z))
DEFOP(op_viaddmnmx16x2, 1, return __viaddmax_s16x2(x, y, z))
DEFOP(op_viadd16x2, 1, return __vadd2(x, y))
DEFOP(op_vabsdiff4, 1, return __vabsdiffu4(x, y))
DEFOP(op_hadd2, 1, return h2u(__hadd2(u2h(x), u2h(y))))
DEFOP(op_hfma2, 1, return h2u(__hfma2(u2h(x), u2h(y), u2h(z))))
DEFOP(op_hmnmx2, 1, return h2u(__hmax2(u2h(x), u2h(y))))
DEFOP(op_hset2, 1, return __hgt2_mask(u2h(x), u2h(y)))
DEFOP(op_fadd, 1, return __float_as_uint(__uint_as_float(x) + __uint_as_float(y)))
DEFOP(op_ffma, 1, return __float_as_uint(__fmaf_rn(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z))))
DEFOP(op_fmnmx, 1, return __float_as_uint(fmaxf(__uint_as_float(x), __uint_as_float(y))))
DEFOP(op_sel, 3, return ((int)x > (int)y) ? z : x + 1)
DEFOP(op_popc, 2, return __popc(x) + y)
DEFOP(op_imad_hi, 1, return __umulhi(x, y) + z)
DEFOP(op_dp4a, 1, return (unsigned)__dp4a((int)x, (int)y, (int)z))
DEFOP(op_dp2a, 1, return (unsigned)__dp2a_lo((int)x, (int)y, (int)z))
DEFOP(op_lea, 1, return (x << 3) + y)
DEFOP(op_shr_imm, 1, return __funnelshift_r(x, y, 12))
DEFOP(op_vabsdiff2, 1, return __vabsdiffs2(x, y))
DEFOP(op_viaddmnmx16x2_relu, 1, return __viaddmin_s16x2_relu(x, y, z))
// mixes (n = instructions per op() call)
DEFOP(mix_vimnmx16_imadhi, 2, return __umulhi(__vmaxs2(x, y), z) + y)
DEFOP(mix_vimnmx16_imad, 2, return __vmaxs2(x, y) * z + y)
DEFOP(mix_alu3_imad1, 4, return __vmaxs2(__vabsdiffu4(__viaddmax_s16x2(x, y, z), y), z) * y + z)
DEFOP(mix_alu2_imad1, 3, return __vmaxs2(__vabsdiffu4(x, y), z) * y + z)
DEFOP(mix_alu2_viadd16_1, 3, return __vadd2(__vmaxs2(__vabsdiffu4(x, y), z), y))
DEFOP(mix_viadd16_hfma2, 2, return h2u(__hfma2(u2h(__vadd2(x, y)), u2h(z), u2h(y))))
DEFOP(mix_viadd16_imad, 2, return __vadd2(x, y) * z + y)
DEFOP(mix_viadd32_imad, 2, return (x + y) * z + y)
DEFOP(mix_alu2_hfma2_2, 4, return h2u(__hfma2(__hsub2_sat(u2h(__vmaxs2(__vabsdiffu4(x, y), z)), u2h(y)), u2h(z), u2h(x))))
DEFOP(mix_alu2_imad1_viadd1, 4, return __vadd2(__vmaxs2(__vabsdiffu4(x, y), z) * y + z, x))
DEFOP(mix_lop3_imad, 2, return ((x & y) ^ z) * y + z)
DEFOP(mix_lop3_hadd2, 2, return h2u(__hadd2(u2h((x & y) ^ z), u2h(y))))
DEFOP(mix_vimnmx16_hadd2, 2, return h2u(__hadd2(u2h(__vmaxs2(x, y)), u2h(z))))
DEFOP(mix_lop3_ffma, 2, return __float_as_uint(__fmaf_rn(__uint_as_float((x & y) ^ z), __uint_as_float(y), __uint_as_float(z))))
DEFOP(mix_lop3_imad_hadd2, 3, return h2u(__hadd2(u2h(((x & y) ^ z) * y + z), u2h(y))))
DEFOP(mix_vimnmx16_viadd16, 2, return __vadd2(__vmaxs2(x, y), z))
DEFOP(mix_hadd2_hmnmx2, 2, return h2u(__hmax2(__hadd2(u2h(x), u2h(y)), u2h(z))))
DEFOP(mix_imad_hadd2, 2, return h2u(__hadd2(u2h(x * y + z), u2h(y))))
DEFOP(mix_lop3_lop3_imad, 3, return (((x & y) ^ z) | (y & 0x55555555u)) * y + z)

template <class OP>
void run(const char* name, int nsm, int threads, int blocks_per_sm) {
    int blocks = nsm * blocks_per_sm;
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * sizeof(long long));
    int iters = 200;
    bench<OP><<<blocks, threads>>>(out, 12345u, 10, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<blocks, threads>>>(out, 12345u, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= blocks;
    double warp_instr_per_sm = (double)iters * INNER * ILP * OP::n * (threads / 32) * blocks_per_sm;
    double ipc = warp_instr_per_sm / avg;
    double total_lane_ops = warp_instr_per_sm * 32.0 * nsm;
    printf("{\"op\": \"%s\", \"threads_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, "
           "\"lane_ops_per_clk_per_sm\": %.1f, \"ms\": %.3f, \"chip_lane_ops_per_s\": %.4e, \"eff_clock_mhz\": %.0f}\n",
           name, threads * blocks_per_sm, ipc, ipc * 32, ms, total_lane_ops / (ms * 1e-3), avg / (ms * 1e-3) / 1e6);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
#define RUN(OP) run<OP>(#OP, nsm, 1024, 1);
    RUN(op_iadd3) RUN(op_lop3) RUN(op_imad) RUN(op_shf) RUN(op_prmt)
    RUN(op_vimnmx32) RUN(op_vimnmx3_32) RUN(op_viaddmnmx32)
    RUN(op_vimnmx16x2) RUN(op_vimnmx3_16x2) RUN(op_viaddmnmx16x2) RUN(op_viadd16x2) RUN(op_vabsdiff4)
    RUN(op_hadd2) RUN(op_hfma2) RUN(op_hmnmx2) RUN(op_hset2)
    RUN(op_fadd) RUN(op_ffma) RUN(op_fmnmx) RUN(op_sel) RUN(op_popc)
    RUN(mix_lop3_imad) RUN(mix_lop3_hadd2) RUN(mix_vimnmx16_hadd2) RUN(mix_lop3_ffma)
    RUN(mix_lop3_imad_hadd2) RUN(mix_vimnmx16_viadd16) RUN(mix_hadd2_hmnmx2) RUN(mix_imad_hadd2)
    RUN(mix_lop3_lop3_imad)
    RUN(op_imad_hi) RUN(op_dp4a) RUN(op_dp2a) RUN(op_lea) RUN(op_shr_imm) RUN(op_vabsdiff2) RUN(op_viaddmnmx16x2_relu)
    RUN(mix_vimnmx16_imadhi) RUN(mix_vimnmx16_imad) RUN(mix_alu3_imad1) RUN(mix_alu2_imad1) RUN(mix_alu2_viadd16_1)
    RUN(mix_alu2_imad1_viadd1)
    RUN(mix_viadd16_hfma2) RUN(mix_viadd16_imad) RUN(mix_viadd32_imad) RUN(mix_alu2_hfma2_2)
    return 0;
}
