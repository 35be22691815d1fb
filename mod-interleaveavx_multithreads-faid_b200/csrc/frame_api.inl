// frame_api.inl -- C-ABI entry points for frame generation, encoding, scoring, the fused Monte-Carlo round and
// the counter all-reduce.  Included at the end of ldpc_b200.cu (shares its helpers and the handle definition).
#include <dlfcn.h>

namespace {

__global__ void generate_bpsk_kernel(const int8_t* __restrict__ output_bits, int8_t* __restrict__ fix, float* __restrict__ sym_out,
                                     int64_t n, float sigma, float scale, int qbits, uint64_t seed, uint64_t first_frame, int tx_reuse,
                                     uint32_t tx_c0, uint64_t group0) {
    // CSimulate.cpp:121-124 + CModulate.cpp:363-370: x = 2b-1 on the two-region buffer, LLR = received amplitude.
    // Position -> (frame, index) only for the Philox counter; 2 normals per call, thread handles an aligned pair.
    // Codeword reuse (CSimulate.cpp:103-117): group g of the launch transmits group (group0 + g) / tx_reuse - tx_c0 of
    // output_bits, exactly as the QPSK / QAM producers do (tx_group_of).
    for (int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; p < n; p += (int64_t)gridDim.x * blockDim.x * 2) {
        const int64_t group = p / (32 * kN);
        const int off = (int)(p - group * 32 * kN);
        int frame, idx;
        if (off < 32 * kK) { frame = off / kK; idx = off - frame * kK; }
        else { frame = (off - 32 * kK) / kM; idx = kK + (off - 32 * kK) - frame * kM; }
        uint32_t r[4];
        Philox::gen(seed, first_frame + (uint64_t)(group * 32 + frame), kNoiseStream + (uint64_t)(idx >> 1), r);
        const float u1 = ((float)r[0] + 0.5f) * 2.3283064365386963e-10f;
        const float u2 = (float)r[1] * 2.3283064365386963e-10f;
        const float rad = sigma * sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        const int64_t txg = tx_reuse > 1 ? (int64_t)((group0 + (uint64_t)group) / (uint64_t)tx_reuse - tx_c0) : group;
        const int64_t ps = txg * 32 * kN + off;
        const float b0 = output_bits ? (float)(2 * output_bits[ps] - 1) : -1.0f;
        const float b1 = output_bits ? (float)(2 * output_bits[ps + 1] - 1) : -1.0f;
        const float y0 = __fadd_rn(b0, __fmul_rn(rad, cs)), y1 = __fadd_rn(b1, __fmul_rn(rad, sn));
        if (sym_out) { sym_out[p] = y0; sym_out[p + 1] = y1; }
        fix[p] = (int8_t)quant_cfg(y0, scale, qbits);
        fix[p + 1] = (int8_t)quant_cfg(y1, scale, qbits);
    }
}

int ensure_tmp(FrameState& fs, int k, size_t bytes) {
    if (fs.tmp_bytes[k] >= bytes) return LDPC_B200_OK;
    if (fs.d_tmp[k]) cudaFree(fs.d_tmp[k]);
    fs.d_tmp[k] = nullptr;
    fs.tmp_bytes[k] = 0;
    CUDA_TRY(cudaMalloc(&fs.d_tmp[k], bytes));
    fs.tmp_bytes[k] = bytes;
    return LDPC_B200_OK;
}

// device view of an input buffer (staged into tmp slot k when it is a host pointer)
int dev_in(ldpc_b200_handle* h, int k, const void* p, size_t bytes, const void** out) {
    if (!p) { *out = nullptr; return LDPC_B200_OK; }
    if (is_device_ptr(p)) { *out = p; return LDPC_B200_OK; }
    int rc = ensure_tmp(h->fs, k, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->fs.d_tmp[k], p, bytes, cudaMemcpyHostToDevice, h->fs.stream));
    *out = h->fs.d_tmp[k];
    return LDPC_B200_OK;
}
// device buffer for an output (tmp slot k when the destination is a host pointer)
int dev_out(ldpc_b200_handle* h, int k, void* p, size_t bytes, void** out) {
    if (!p) { *out = nullptr; return LDPC_B200_OK; }
    if (is_device_ptr(p)) { *out = p; return LDPC_B200_OK; }
    int rc = ensure_tmp(h->fs, k, bytes);
    if (rc) return rc;
    *out = h->fs.d_tmp[k];
    return LDPC_B200_OK;
}
int copy_back(ldpc_b200_handle* h, void* host, const void* dev, size_t bytes) {
    if (!host || host == dev) return LDPC_B200_OK;
    CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, h->fs.stream));
    return LDPC_B200_OK;
}

// sigma of CSimulate::Configure (CSimulate.cpp:67-75) and what CSimulate::Run passes to the channel (:126)
float sim_sigma(const ldpc_b200_config& c, float ebn0) {
    if (c.mod_type == 1) return (float)(1.0 / sqrt(2.0 * c.code_rate * c.mod_type * pow(10.0, 0.1 * ebn0)));
    return (float)(1.0 / sqrt(c.code_rate * c.mod_type * pow(10.0, 0.1 * ebn0)));
}

int grid_for(int64_t items, int block) { return (int)std::min<int64_t>((items + block - 1) / block, 148 * 16); }

// identity interleaver and word-aligned bit buffers: the producer may fetch the transmitted bits 32 bits at a time
int gen_fast_i1(const ldpc_b200_config& c, const int8_t* d_tx, const int8_t* d_codeword) {
    return c.interleave_mod_type == 1 && c.mod_type != 1 && ((uintptr_t)d_tx & 3) == 0 && ((uintptr_t)d_codeword & 3) == 0;
}

int launch_generate(ldpc_b200_handle* h, const int8_t* d_tx, const int8_t* d_codeword, const float* d_sym_in,
                    float* d_sym_out, float* d_llr, int8_t* d_fix, int n_groups, float ebn0, uint64_t seed,
                    uint64_t first_frame, bool add_noise, int tx_reuse = 1, uint32_t tx_c0 = 0, uint64_t group0 = 0) {
    const ldpc_b200_config& c = h->cfg;
    cudaStream_t st = h->fs.stream;
    if (c.mod_type == 1) {
        if (d_sym_in || d_llr) return fail(LDPC_B200_EINVAL, "BPSK: demap / float LLR outputs are not defined (the reference quantises the received amplitude directly)");
        const int64_t n = (int64_t)n_groups * 32 * kN;
        const int8_t* tx = d_tx;
        if (d_codeword) return fail(LDPC_B200_EINVAL, "BPSK generate needs outputBits (or NULL for the all-zero codeword)");
        generate_bpsk_kernel<<<grid_for(n / 2, 256), 256, 0, st>>>(tx, d_fix, d_sym_out, n, sim_sigma(c, ebn0), c.scale, c.quant_bits ? c.quant_bits : 4, seed, first_frame,
                                                                   tx_reuse, tx_c0, group0);
        CUDA_TRY(cudaGetLastError());
        return LDPC_B200_OK;
    }
    GenParams P;
    memset(&P, 0, sizeof P);
    P.core.output_bits = d_tx;
    P.core.codeword = d_codeword;
    P.symbols_in = d_sym_in;
    P.symbols_out = d_sym_out;
    P.llr_float = d_llr;
    P.fix = d_fix;
    P.n_groups = n_groups;
    P.core.mod = c.mod_type;
    P.core.I = c.interleave_mod_type;
    P.core.sigma_d = (float)(sim_sigma(c, ebn0) / sqrt(2));
    P.core.scale = c.scale;
    P.core.qbits = c.quant_bits ? c.quant_bits : 4;
    P.core.seed = seed;
    P.core.first_frame = first_frame;
    P.core.add_noise = add_noise ? 1 : 0;
    const int64_t npairs = (int64_t)n_groups * 32 * kN / c.mod_type / 2;
    P.core.fast_i1 = gen_fast_i1(c, d_tx, d_codeword);
    P.core.tx_reuse = tx_reuse;
    P.core.tx_c0 = tx_c0;
    P.core.group0 = group0;
    if (P.core.fast_i1 && add_noise && !d_sym_in && !d_sym_out && !d_llr && d_fix && ((uintptr_t)d_fix & 3) == 0) {
        if (c.mod_type == 2) generate_i1_kernel<2><<<grid_for(npairs, 256), 256, 0, st>>>(P);
        else if (c.mod_type == 4) generate_i1_kernel<4><<<grid_for(npairs, 256), 256, 0, st>>>(P);
        else if (c.mod_type == 6) generate_i1_kernel<6><<<grid_for(npairs, 256), 256, 0, st>>>(P);
        else generate_i1_kernel<8><<<grid_for(npairs, 256), 256, 0, st>>>(P);
    } else {
        generate_kernel<<<grid_for(npairs, 256), 256, 0, st>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    return LDPC_B200_OK;
}

int launch_encode(ldpc_b200_handle* h, const int8_t* d_info, int8_t* d_tx, int n_groups) {
    static std::atomic<bool> attr[64];  // per device; handles may be driven from several host threads
    const int dev = h->cfg.device;
    const size_t smem = (size_t)(kK + kM) * sizeof(uint32_t);
    if (dev < 0 || dev >= 64 || !attr[dev].load(std::memory_order_acquire)) {
        CUDA_TRY(cudaFuncSetAttribute(encode_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr[dev].store(true, std::memory_order_release);
    }
    // few groups: split each group over 12 CTAs (one per parity block row) to cut the latency of a lone encode
    encode_group_kernel<<<dim3(n_groups, n_groups <= 64 ? LDPC_MB : 1), kEncThreads, smem, h->fs.stream>>>(d_info, d_tx, n_groups);
    CUDA_TRY(cudaGetLastError());
    return LDPC_B200_OK;
}

// ---- NCCL through dlopen: the engine has no link-time dependency on it ----
typedef struct { char b[128]; } NcclId;
typedef int (*nccl_get_id_t)(NcclId*);
typedef int (*nccl_init_rank_t)(void**, int, NcclId, int);
typedef int (*nccl_allreduce_t)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_destroy_t)(void*);
typedef const char* (*nccl_errstr_t)(int);

void* nccl_handle() {
    static void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // C++11 static initialisation: thread-safe, once
    return lib;
}
template <class T>
T nccl_sym(const char* name) {
    void* lib = nccl_handle();
    return lib ? (T)dlsym(lib, name) : nullptr;
}

void comm_destroy(FrameState& fs) {
    if (!fs.nccl_comm) return;
    if (auto f = nccl_sym<nccl_destroy_t>("ncclCommDestroy")) f(fs.nccl_comm);
    fs.nccl_comm = nullptr;
}

// Caller-supplied DEVICE pointers are read / written with vector accesses; a misaligned torch view would fault and kill the
// context, so it is rejected here instead.  (Host pointers are staged through cudaMalloc'd buffers and need no alignment.)
int check_align(const void* p, size_t a, const char* what) {
    if (p && ((uintptr_t)p & (a - 1))) return fail(LDPC_B200_EINVAL, std::string(what) + " must be " + std::to_string(a) + "-byte aligned");
    return LDPC_B200_OK;
}

}  // namespace

extern "C" {

int ldpc_b200_quantize(ldpc_b200_handle* h, const float* in, int8_t* out, int64_t length, float scale) {
    return ldpc_b200_quantize_bits(h, in, out, length, scale, 4);
}

int ldpc_b200_quantize_bits(ldpc_b200_handle* h, const float* in, int8_t* out, int64_t length, float scale, int bits) {
    if (!h || !in || !out || length < 0) return fail(LDPC_B200_EINVAL, "quantize: bad arguments");
    if (bits < 1 || bits > 6) return fail(LDPC_B200_EINVAL, "quantize: bits must be 1..6");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (length == 0) return LDPC_B200_OK;
    const void* din; void* dout;
    int rc = dev_in(h, 0, in, (size_t)length * 4, &din);
    if (rc) return rc;
    rc = dev_out(h, 1, out, (size_t)length, &dout);
    if (rc) return rc;
    quantize_kernel<<<grid_for(length, 256), 256, 0, h->fs.stream>>>((const float*)din, (int8_t*)dout, length, scale, bits);
    CUDA_TRY(cudaGetLastError());
    rc = copy_back(h, out, dout, (size_t)length);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    return LDPC_B200_OK;
}

int ldpc_b200_demap(ldpc_b200_handle* h, const float* symbols, int n_groups, float* llr_float, int8_t* fixInput) {
    if (!h || !symbols || n_groups < 0 || (!llr_float && !fixInput)) return fail(LDPC_B200_EINVAL, "demap: bad arguments");
    if (h->cfg.mod_type == 1) return fail(LDPC_B200_EINVAL, "demap: BPSK has no demapper in the reference (CSimulate.cpp:121-124)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    const size_t nsym = (size_t)n_groups * 32 * kN / h->cfg.mod_type, nb = (size_t)n_groups * 32 * kN;
    const void* dsym; void *dllr, *dfix;
    int rc = dev_in(h, 0, symbols, nsym * 8, &dsym);
    if (!rc) rc = dev_out(h, 1, llr_float, nb * 4, &dllr);
    if (!rc) rc = dev_out(h, 2, fixInput, nb, &dfix);
    if (!rc) rc = check_align(dsym, 16, "demap: symbols");
    if (rc) return rc;
    rc = launch_generate(h, nullptr, nullptr, (const float*)dsym, nullptr, (float*)dllr, (int8_t*)dfix, n_groups, 0.f, 0, 0, false);
    if (rc) return rc;
    if (!(rc = copy_back(h, llr_float, dllr, nb * 4))) rc = copy_back(h, fixInput, dfix, nb);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    return LDPC_B200_OK;
}

int ldpc_b200_generate(ldpc_b200_handle* h, const int8_t* outputBits, float ebn0_db, uint64_t seed, uint64_t first_frame_index,
                       int n_groups, float* symbols_out, int8_t* fixInput) {
    if (!h || n_groups < 0 || !fixInput) return fail(LDPC_B200_EINVAL, "generate: bad arguments");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    const size_t nb = (size_t)n_groups * 32 * kN;
    const size_t nsym_f = h->cfg.mod_type == 1 ? nb : 2 * nb / h->cfg.mod_type;  // floats
    const void* dtx; void *dsym, *dfix;
    int rc = dev_in(h, 0, outputBits, nb, &dtx);
    if (!rc) rc = dev_out(h, 1, symbols_out, nsym_f * 4, &dsym);
    if (!rc) rc = dev_out(h, 2, fixInput, nb, &dfix);
    if (!rc && h->cfg.mod_type != 1) rc = check_align(dsym, 16, "generate: symbols_out");
    if (rc) return rc;
    const int8_t* cw = nullptr;
    if (!outputBits && h->cfg.mod_type != 1) {  // the shipped CodeWord_sym is all-zero (Codeword.h:4)
        CUDA_TRY(cudaMemsetAsync(h->fs.d_codeword, 0, kN, h->fs.stream));
        cw = h->fs.d_codeword;
    }
    rc = launch_generate(h, (const int8_t*)dtx, cw, nullptr, (float*)dsym, nullptr, (int8_t*)dfix, n_groups, ebn0_db, seed,
                         first_frame_index, true);
    if (rc) return rc;
    if (!(rc = copy_back(h, symbols_out, dsym, nsym_f * 4))) rc = copy_back(h, fixInput, dfix, nb);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    return LDPC_B200_OK;
}

int ldpc_b200_gen_msg_seq(ldpc_b200_handle* h, uint64_t seed, uint64_t first_frame_index, int n_groups, int8_t* inputBits) {
    if (!h || !inputBits || n_groups < 0) return fail(LDPC_B200_EINVAL, "gen_msg_seq: bad arguments");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    const size_t nb = (size_t)n_groups * 32 * kK;
    void* dout;
    int rc = dev_out(h, 0, inputBits, nb, &dout);
    if (!rc) rc = check_align(dout, 4, "gen_msg_seq: inputBits");
    if (rc) return rc;
    info_bits_kernel<<<grid_for((int64_t)n_groups * 32 * (kK / 128), 256), 256, 0, h->fs.stream>>>((int8_t*)dout, n_groups, seed, first_frame_index, 1);
    CUDA_TRY(cudaGetLastError());
    rc = copy_back(h, inputBits, dout, nb);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    return LDPC_B200_OK;
}

int ldpc_b200_encode(ldpc_b200_handle* h, const int8_t* inputBits, int8_t* outputBits, int n_groups) {
    if (!h || !inputBits || !outputBits || n_groups < 0) return fail(LDPC_B200_EINVAL, "encode: bad arguments");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    const void* din; void* dout;
    int rc = dev_in(h, 0, inputBits, (size_t)n_groups * 32 * kK, &din);
    if (!rc) rc = dev_out(h, 1, outputBits, (size_t)n_groups * 32 * kN, &dout);
    if (rc) return rc;
    rc = launch_encode(h, (const int8_t*)din, (int8_t*)dout, n_groups);
    if (rc) return rc;
    rc = copy_back(h, outputBits, dout, (size_t)n_groups * 32 * kN);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    return LDPC_B200_OK;
}

int ldpc_b200_count_errors(ldpc_b200_handle* h, const int8_t* inputBits, const int8_t* decodedBits, int n_groups,
                           uint64_t* counters) {
    if (!h || !inputBits || !decodedBits || !counters || n_groups < 0) return fail(LDPC_B200_EINVAL, "count_errors: bad arguments");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    const void *din, *ddec;
    int rc = dev_in(h, 0, inputBits, (size_t)n_groups * 32 * kK, &din);
    if (!rc) rc = dev_in(h, 1, decodedBits, (size_t)n_groups * 32 * kN, &ddec);
    if (!rc) rc = check_align(din, 16, "count_errors: inputBits");
    if (!rc) rc = check_align(ddec, 16, "count_errors: decodedBits");
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(h->fs.d_counters, 0, LDPC_B200_NUM_COUNTERS * 8, h->fs.stream));
    const int frames = n_groups * 32;
    count_errors_kernel<<<std::min((frames + 7) / 8, 148 * 8), 256, 0, h->fs.stream>>>((const int8_t*)din, (const int8_t*)ddec, frames, h->fs.d_counters, kK, 1, 0u, 0ull);
    CUDA_TRY(cudaGetLastError());
    uint64_t tmp[LDPC_B200_NUM_COUNTERS];
    CUDA_TRY(cudaMemcpyAsync(tmp, h->fs.d_counters, sizeof tmp, cudaMemcpyDeviceToHost, h->fs.stream));
    CUDA_TRY(cudaStreamSynchronize(h->fs.stream));
    for (int i = 0; i < LDPC_B200_NUM_COUNTERS; ++i) counters[i] += tmp[i];
    return LDPC_B200_OK;
}

int ldpc_b200_simulate(ldpc_b200_handle* h, const int8_t* codeword, float ebn0_db, uint64_t seed, uint64_t first_frame_index,
                       int n_groups, uint64_t* counters) {
    if (!h || !counters || n_groups < 0) return fail(LDPC_B200_EINVAL, "simulate: bad arguments");
    if (!codeword && first_frame_index % 32)
        return fail(LDPC_B200_EINVAL, "simulate: with random info bits first_frame_index must be a multiple of 32 (codeword reuse is per whole group)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (n_groups == 0) return LDPC_B200_OK;
    FrameState& fs = h->fs;
    const int cg = std::min(h->chunk_groups, 512);
    if (fs.sim_groups < cg) {
        if (fs.d_info) cudaFree(fs.d_info);
        if (fs.d_tx) cudaFree(fs.d_tx);
        if (fs.d_fix) cudaFree(fs.d_fix);
        if (fs.d_dec) cudaFree(fs.d_dec);
        fs.d_info = fs.d_tx = fs.d_fix = fs.d_dec = nullptr;
        fs.sim_groups = 0;
        CUDA_TRY(cudaMalloc(&fs.d_info, (size_t)cg * 32 * kK));
        CUDA_TRY(cudaMalloc(&fs.d_tx, (size_t)cg * 32 * kN));
        CUDA_TRY(cudaMalloc(&fs.d_fix, (size_t)cg * 32 * kN));
        CUDA_TRY(cudaMalloc(&fs.d_dec, (size_t)cg * 32 * kN));
        fs.sim_groups = cg;
    }
    h->last_kernel_ms = h->last_decode_ms = h->last_finalize_ms = 0.f;
    h->last_launches = 0;
    Slot& s = h->slots[0];
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    // BPSK keeps the separate producer kernel (its real-valued channel has its own kernel); the environment switch is
    // for A/B measurements and for the test that both paths give identical counters
    const bool fused = h->cfg.mod_type != 1 && getenv("LDPC_B200_NO_FUSED_PRODUCER") == nullptr;
    // everything of one round runs in order on the decode slot's stream
    cudaStream_t saved = fs.stream;
    fs.stream = s.stream;
    int rc = LDPC_B200_OK;
    do {
        if ((rc = cudaMemsetAsync(fs.d_counters, 0, LDPC_B200_NUM_COUNTERS * 8, s.stream)) != cudaSuccess) { rc = fail(LDPC_B200_ECUDA, "memset counters"); break; }
        if (codeword) {
            if (cudaMemcpyAsync(fs.d_codeword, codeword, kN, cudaMemcpyDefault, s.stream) != cudaSuccess) { rc = fail(LDPC_B200_ECUDA, "copy codeword"); break; }
        }
        for (int g0 = 0; g0 < n_groups && !rc; g0 += cg) {
            const int groups = std::min(cg, n_groups - g0);
            const uint64_t ff = first_frame_index + (uint64_t)g0 * 32;
            const int8_t* d_tx = nullptr;
            // codeword reuse (CSimulate.cpp:103-117): the groups of one "run" of `reuse` consecutive global groups share
            // their 32 encoded frames; only the noise differs
            const int reuse = std::max(1, h->cfg.codeword_reuse ? h->cfg.codeword_reuse : 50);
            const uint64_t G0 = ff / 32;
            const uint64_t c0 = G0 / (uint64_t)reuse, c1 = (G0 + (uint64_t)groups - 1) / (uint64_t)reuse;
            const int ncw = (int)(c1 - c0 + 1);
            if (!codeword) {
                info_bits_kernel<<<grid_for((int64_t)ncw * 32 * (kK / 128), 256), 256, 0, s.stream>>>(fs.d_info, ncw, seed, c0 * (uint64_t)reuse * 32, reuse);
                if ((rc = launch_encode(h, fs.d_info, fs.d_tx, ncw))) break;
                d_tx = fs.d_tx;
                h->last_launches += 2;
            } else if (h->cfg.mod_type == 1) {
                rc = fail(LDPC_B200_EINVAL, "simulate: BPSK with a fixed codeword is not supported; pass codeword = NULL");
                break;
            }
            if (fused) {
                // producer fused into the decoder's loader: no LLR buffer, one launch less
                GenCore G;
                memset(&G, 0, sizeof G);
                G.output_bits = d_tx;
                G.codeword = codeword ? fs.d_codeword : nullptr;
                G.mod = h->cfg.mod_type;
                G.I = h->cfg.interleave_mod_type;
                G.sigma_d = (float)(sim_sigma(h->cfg, ebn0_db) / sqrt(2));
                G.scale = h->cfg.scale;
                G.qbits = h->cfg.quant_bits ? h->cfg.quant_bits : 4;
                G.seed = seed;
                G.first_frame = ff;
                G.add_noise = 1;
                G.fast_i1 = gen_fast_i1(h->cfg, G.output_bits, G.codeword);
                G.tx_reuse = codeword ? 1 : reuse;
                G.tx_c0 = (uint32_t)c0;
                G.group0 = G0;
                if ((rc = run_chunk(h, s, nullptr, false, fs.d_dec, nullptr, groups, &G))) break;
            } else {
                if ((rc = launch_generate(h, d_tx, codeword ? fs.d_codeword : nullptr, nullptr, nullptr, nullptr, fs.d_fix, groups,
                                          ebn0_db, seed, ff, true, codeword ? 1 : reuse, (uint32_t)c0, G0))) break;
                h->last_launches += 1;
                if ((rc = run_chunk(h, s, fs.d_fix, false, fs.d_dec, nullptr, groups))) break;
            }
            const int frames = groups * 32;
            count_errors_kernel<<<std::min((frames + 7) / 8, 148 * 8), 256, 0, s.stream>>>(
                codeword ? fs.d_codeword : fs.d_info, fs.d_dec, frames, fs.d_counters, codeword ? 0 : kK, codeword ? 1 : reuse,
                (uint32_t)c0, G0);
            group_hist_kernel<<<(groups + 255) / 256, 256, 0, s.stream>>>(s.d_bf, s.d_its, groups, fs.d_counters);
            h->last_launches += 2;
            if (cudaGetLastError() != cudaSuccess) { rc = fail(LDPC_B200_ECUDA, "simulate launch"); break; }
            if ((rc = collect_timing(h, s))) break;
        }
    } while (0);
    fs.stream = saved;
    if (rc) return rc;
    uint64_t tmp[LDPC_B200_NUM_COUNTERS];
    CUDA_TRY(cudaMemcpyAsync(tmp, fs.d_counters, sizeof tmp, cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < LDPC_B200_NUM_COUNTERS; ++i) counters[i] += tmp[i];
    return LDPC_B200_OK;
}

int ldpc_b200_nccl_unique_id(uint8_t unique_id[128]) {
    auto f = nccl_sym<nccl_get_id_t>("ncclGetUniqueId");
    if (!f) return fail(LDPC_B200_ENCCL, "libnccl.so.2 not found");
    NcclId id;
    int e = f(&id);
    if (e) return fail(LDPC_B200_ENCCL, "ncclGetUniqueId failed");
    memcpy(unique_id, id.b, 128);
    return LDPC_B200_OK;
}

int ldpc_b200_comm_init(ldpc_b200_handle* h, const uint8_t unique_id[128], int rank, int n_ranks) {
    if (!h || !unique_id || rank < 0 || rank >= n_ranks) return fail(LDPC_B200_EINVAL, "comm_init: bad arguments");
    auto f = nccl_sym<nccl_init_rank_t>("ncclCommInitRank");
    if (!f) return fail(LDPC_B200_ENCCL, "libnccl.so.2 not found");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    comm_destroy(h->fs);  // a second comm_init replaces the communicator (e.g. the job was re-sharded)
    NcclId id;
    memcpy(id.b, unique_id, 128);
    int e = f(&h->fs.nccl_comm, n_ranks, id, rank);
    if (e) {
        auto es = nccl_sym<nccl_errstr_t>("ncclGetErrorString");
        return fail(LDPC_B200_ENCCL, std::string("ncclCommInitRank: ") + (es ? es(e) : "error"));
    }
    return LDPC_B200_OK;
}

int ldpc_b200_allreduce_counters(ldpc_b200_handle* h, uint64_t* counters) {
    if (!h || !counters) return fail(LDPC_B200_EINVAL, "allreduce_counters: bad arguments");
    if (!h->fs.nccl_comm) return fail(LDPC_B200_ENCCL, "allreduce_counters: call ldpc_b200_comm_init first");
    auto f = nccl_sym<nccl_allreduce_t>("ncclAllReduce");
    if (!f) return fail(LDPC_B200_ENCCL, "libnccl.so.2 not found");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    cudaStream_t st = h->fs.stream;
    CUDA_TRY(cudaMemcpyAsync(h->fs.d_counters, counters, LDPC_B200_NUM_COUNTERS * 8, cudaMemcpyHostToDevice, st));
    int e = f(h->fs.d_counters, h->fs.d_counters, LDPC_B200_NUM_COUNTERS, /*ncclUint64*/ 5, /*ncclSum*/ 0, h->fs.nccl_comm, st);
    if (e) return fail(LDPC_B200_ENCCL, "ncclAllReduce failed");
    CUDA_TRY(cudaMemcpyAsync(counters, h->fs.d_counters, LDPC_B200_NUM_COUNTERS * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return LDPC_B200_OK;
}

}  // extern "C"
