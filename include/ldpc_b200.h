/* ldpc_b200.h -- C-ABI of the B200-native LDPC decoding engine (libldpc_b200.so).
 *
 * Drop-in boundary for the hot path of Lcrypto/mod-interleaveavx_multithreads-FAID: everything
 * CSimulate::Run (CSimulate.cpp:92-180) does per 32-frame block between "transmitted bits" and
 * "error counters".  Plain pointers and sizes only; no C++/torch types.  All functions return 0 on
 * success or a negative LDPC_B200_E* code -- they never exit() or block on stdin (the reference does,
 * CTool.cpp:591-596).  A handle is NOT thread-safe: one handle per host thread / GPU, mirroring the
 * reference's one-object-set-per-pthread rule (CSimulate.cpp:218-278).
 *
 * Reference interface each entry point replaces:
 *   ldpc_b200_read_profile      ReadProfile(Parameter_Simulation*)            CTool.cpp:588-621
 *   ldpc_b200_default_config    compile-time constants                        CDecoder_FAID.cpp:4-170, CDecoder_FAID_2B1C.cpp:6-90,
 *                                                                             CDecoder_OMS*.cpp:1-36, CTool.h:6, Constants_SSE.h:20-25
 *   ldpc_b200_create/destroy    new CLDPC + CLDPC::Initial / ~CLDPC           CLDPC.cpp:17-55,4772-4817; CSimulate.cpp:41-59
 *   ldpc_b200_decode            CLDPC::Decode, Decode_OMS, Decode_FAID,       CLDPC.cpp:214, CDecoder_OMS.cpp:13, CDecoder_FAID.cpp:176,
 *                               Decode_OMSBF, Decode_OMS_DTBF,                CDecoder_OMSBF.cpp:13, CDecoder_OMS_DTBF.cpp:18,
 *                               Decode_FAID_2B1C  (fixInput -> decodedBits)   CDecoder_FAID_2B1C.cpp:96; dispatch CSimulate.cpp:136-164
 *   ldpc_b200_quantize          CLDPC::float2LimitChar_4bit                   CLDPC.cpp:4524-4582
 *   ldpc_b200_quantize_bits     CLDPC::float2LimitChar_{1,2,3,5,6}bit         CLDPC.cpp:4385-4522,4584-4770
 *   ldpc_b200_demap             CModulate::Demodulation + AfterDeModulationDeInterleaver + float2LimitChar_4bit
 *                                                                             CModulate.cpp:156-212,270-362; CSimulate.cpp:127-132
 *   ldpc_b200_generate          CModulate::BeforeModulationInterleaver/Modulation + CChannel::AWGNChannel + the three above
 *                                                                             CModulate.cpp:95-152,216-264; CChannel.cpp:71-97; CSimulate.cpp:111-132
 *   ldpc_b200_gen_msg_seq       CLDPC::GenMsgSeq                              CLDPC.cpp:60-66
 *   ldpc_b200_encode            CLDPC::Encode / FakeEncoder                   CLDPC.cpp:68-207
 *   ldpc_b200_count_errors      CLDPC::CalculateErrors                        CLDPC.cpp:4819-4995
 *   ldpc_b200_simulate          CSimulate::Run (the 50-block frame loop)      CSimulate.cpp:92-180
 *   ldpc_b200_allreduce_counters  the join-and-sum over threads               main.cpp:170-182
 *
 * Buffer layouts are the reference's, per group of 32 frames:
 *   fixInput    int8[32*N]  info region [f][j] at f*K+j (j<K), parity region at 32*K + f*M + j.  Values are in [-7,7] after
 *               float2LimitChar_4bit; ANY int8 value is accepted and decoded exactly as the reference's 8-bit saturating
 *               arithmetic would (a first V2C is clamped at -31 by every decoder and, by the FAID decoders, at +31; the
 *               min-sum decoders leave it unclamped above, where every L >= 39 acts like 39 -- CLDPC.cpp:330,390-397;
 *               pinned against the compiled reference on full-range inputs, tests/test_oracle_vs_reference.py)
 *   decodedBits int8[32*N]  values 0/1, frame-major f*N + n (whole codeword)
 *   inputBits   int8[32*K]  info bits, frame-major
 *   outputBits  int8[32*N]  transmitted bits, same two-region layout as fixInput
 * Group g of a multi-group call starts at g*32*N (resp. g*32*K).
 * Pointers may be host or device pointers (detected with cudaPointerGetAttributes); host buffers are
 * staged through pinned memory in chunks, overlapped with the kernels.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define LDPC_B200_API __attribute__((visibility("default")))
#else
#define LDPC_B200_API
#endif

#define LDPC_B200_N 17664
#define LDPC_B200_M 3072
#define LDPC_B200_K 14592
#define LDPC_B200_GROUP 32 /* frames per decode call of the reference ("group") */

/* error codes */
#define LDPC_B200_OK 0
#define LDPC_B200_EINVAL -1   /* bad argument / unsupported configuration */
#define LDPC_B200_ECUDA -2    /* CUDA runtime error (see ldpc_b200_last_error) */
#define LDPC_B200_ENOMEM -3
#define LDPC_B200_EIO -4      /* Profile.txt could not be read / parsed */
#define LDPC_B200_ENODEV -5   /* no CUDA device: there is NO CPU fallback */
#define LDPC_B200_ENCCL -6

/* DecodeMethod values of Profile.txt (README.md:13-14, CSimulate.cpp:136-164) */
#define LDPC_B200_NMS 0
#define LDPC_B200_OMS 1
#define LDPC_B200_FAID_DTBF 2
#define LDPC_B200_OMS_BF 3
#define LDPC_B200_OMS_DTBF 4
#define LDPC_B200_FAID_2B1C 5

/* LUT variants, selected by #define in the reference (CDecoder_FAID.cpp:8, CDecoder_FAID_2B1C.cpp:12-47) */
#define LDPC_B200_LUT_FAID3 0
#define LDPC_B200_LUT_FAID32 1
#define LDPC_B200_LUT_FAID2 2
#define LDPC_B200_LUT_HYBRID 3

/* BF flavours of the post-processing stage */
#define LDPC_B200_BF_NONE 0
#define LDPC_B200_BF_PLAIN 1 /* CDecoder_OMSBF.cpp:2959-3511 */
#define LDPC_B200_BF_DTBF 2  /* CDecoder_FAID.cpp:6411-7088, CDecoder_OMS_DTBF.cpp:2968-3650 */
#define LDPC_B200_BF_2B1C 3  /* CDecoder_FAID_2B1C.cpp:6124-6813 */

/* By-value configuration: every Profile.txt field plus the reference's compile-time constants. */
typedef struct ldpc_b200_config {
    uint32_t struct_size; /* = sizeof(ldpc_b200_config); checked by create() */
    uint32_t abi_version; /* = LDPC_B200_ABI_VERSION */

    /* --- Profile.txt, in file order (CTool.cpp:597-616) --- */
    float snr_start;             /* StartSNR */
    float snr_pass;              /* SNRPass */
    float snr_end;               /* EndSNR */
    int32_t decode_method;       /* DecodeMethod 0..5 (any other value behaves as 0, CSimulate.cpp:161-163) */
    int32_t max_iteration;       /* MaxIteration: 0..1000 (the reference has no bound; scratch grows with it: one 2.2 KB snapshot per
                                    frame and iteration for the early-stopping methods) */
    int32_t mod_type;            /* modType: 1 BPSK, 2 QPSK, 4 16-QAM, 6 64-QAM, 8 256-QAM */
    int32_t interleave_mod_type; /* InterleaveModType */
    int32_t factor_1;            /* Factor_1 */
    int32_t factor_2;            /* Factor_2 */
    int32_t nb_frames;           /* noFrames: must be 32 */
    float scale;                 /* scale of the 4-bit quantiser */
    int32_t Z;                   /* Z: 256 */

    /* --- compile-time constants of the reference, as data --- */
    int8_t v2c_lut[6][4][8];     /* V2C_map_it{1..6}_[weight class][min(|v|,7)]  (CDecoder_FAID.cpp:12-127) */
    int8_t v2c_lut_ef[6][4][8];  /* V2C_map_it{1..6}_ef                           (CDecoder_FAID.cpp:130-165) */
    int32_t ef_elimination;      /* EF_ELIMINATION 0|1|2 (CDecoder_FAID.cpp:6; 2 = erasure mode, :673-680, DecodeMethod 2 only) */
    int32_t ef_floor_err_count;  /* lane uses the EF LUT iff error_sum < this (signed-saturating count) */
    int32_t ef_floor_iter_thresh;/* ... and remaining iterations <= this */
    int32_t oms_floor_err_count; /* 100 (CDecoder_OMS.cpp:26) */
    int32_t oms_floor_iter_thresh; /* 4 (CDecoder_OMS.cpp:27) */
    int32_t bf_mode;             /* LDPC_B200_BF_* */
    int32_t bf_max_iter;         /* _maxBFiter */
    int32_t dtbf_L0, dtbf_L1, dtbf_delta, dtbf_alpha; /* CDecoder_FAID.cpp:167-170 etc. */
    int32_t regular_col_weight;  /* REGULAR_COL_WEIGHT 3 (CTool.h:6) */
    int32_t hard2_threshold;     /* 13: |L| >= 13 sets the second bit of the 2B1C state (CDecoder_FAID_2B1C.cpp:6130) */
    int32_t puncture_tail;       /* 384 trailing code bits get APP 0 at load (CLDPC.cpp:270-272) */
    double code_rate;            /* 0.8444444 hard-coded (CLDPC.cpp:4780) */

    /* --- execution --- */
    int32_t device;              /* CUDA device ordinal */
    int32_t n_streams;           /* streams used to overlap staging and kernels for host buffers (>=1) */
    int32_t chunk_groups;        /* groups per chunk (0 = library default: 1024 for device-resident buffers; host arrays: 128 when staged, 32 when copied as they are) */
    int32_t quant_bits;          /* LLR quantiser of the producer / demapper: 0 or 4 = float2LimitChar_4bit (the one CSimulate
                                    calls, CSimulate.cpp:124,132); 1,2,3,5,6 = the other float2LimitChar_*bit (CLDPC.cpp:4385-4770) */
    int32_t oms_mode;            /* OMS_MODE of the OMS family (CDecoder_OMS.cpp:3): 1 = selective offset (shipped), 0 = simple:
                                    cste = min(sat8(min - oms_offset), 7) */
    int32_t oms_offset;          /* `offset` of the simple mode (CDecoder_OMS.cpp:6: 1) */
    int32_t codeword_reuse;      /* ldpc_b200_simulate with random info bits: consecutive groups that share one encoded group
                                    (CSimulate::Run encodes once per 50 noise blocks, CSimulate.cpp:103-117).  0 = 50; 1 = fresh bits per group */
    int32_t reserved[1];
} ldpc_b200_config;

typedef struct ldpc_b200_handle ldpc_b200_handle;

/* Counters of one simulation round; index = position in the uint64 vector that is all-reduced.
 * First four mirror CSimulate's public counters (CSimulate.h:36-43). */
enum {
    LDPC_B200_CNT_TEST_FRAME = 0,
    LDPC_B200_CNT_ERROR_FRAME = 1,
    LDPC_B200_CNT_ERROR_BITS = 2,
    LDPC_B200_CNT_LT3_ERR_BIT_FRAME = 3,
    LDPC_B200_CNT_GROUPS = 4,
    LDPC_B200_CNT_MS_ITERS_SUM = 5,   /* sum over groups of executed min-sum iterations */
    LDPC_B200_CNT_BF_HIST = 8,        /* [8 .. 8+51): histogram of BF iterations per group (iterCount.txt) */
    LDPC_B200_CNT_MS_HIST = 64,       /* [64 .. 64+64): histogram of executed min-sum iterations per group */
    LDPC_B200_NUM_COUNTERS = 128
};

LDPC_B200_API const char* ldpc_b200_version(void);
/* Thread-local description of the last failure of a call made from this thread. */
LDPC_B200_API const char* ldpc_b200_last_error(void);

/* Fills *cfg with the reference's shipped constants for `decode_method`, the LUT variant
 * (LDPC_B200_LUT_*; ignored by methods 0,1,3,4; method 5 normally uses LUT_HYBRID) and the shipped
 * Profile.txt values (MaxIteration 6, QPSK, Factor 1/6 -- 26/26 for NMS, scale 13 -- 12.5 for method 5). */
LDPC_B200_API int ldpc_b200_default_config(ldpc_b200_config* cfg, int decode_method, int lut_variant);

/* Positional token parser with the exact token order of ReadProfile (CTool.cpp:597-616).  Overwrites the
 * Profile.txt fields of *cfg and re-derives the method-dependent constants when DecodeMethod changes. */
LDPC_B200_API int ldpc_b200_read_profile(const char* path, ldpc_b200_config* cfg, int lut_variant);

LDPC_B200_API int ldpc_b200_create(const ldpc_b200_config* cfg, ldpc_b200_handle** out);
LDPC_B200_API int ldpc_b200_destroy(ldpc_b200_handle* h);

/* The reference re-reads Factor_1/Factor_2 from Profile.txt inside every decode call; this is the
 * explicit equivalent. */
LDPC_B200_API int ldpc_b200_set_factors(ldpc_b200_handle* h, int factor_1, int factor_2);
LDPC_B200_API int ldpc_b200_set_max_iteration(ldpc_b200_handle* h, int max_iteration);

/* Decode n_groups groups of 32 frames.  Optional outputs (may be NULL), one entry per group unless noted:
 *   bf_iters        BF iterations executed (the int returned by Decode_OMSBF / Decode_OMS_DTBF; also
 *                   reported for methods 2 and 5; 0 for methods 0 and 1)
 *   its_per_group   min-sum iterations executed by the group (MaxIteration unless the group stopped early)
 *   conv_iter       per FRAME (32*n_groups entries): first iteration index (0-based count of completed
 *                   iterations) at whose start the frame's syndrome was zero, or -1 if never observed;
 *                   always -1 for method 0, which has no syndrome check. */
LDPC_B200_API int ldpc_b200_decode(ldpc_b200_handle* h, const int8_t* fixInput, int8_t* decodedBits, int n_groups,
                     int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter);

/* Throughput variant with the engine's native layouts (device or host pointers):
 *   llr_packed   uint8[n_groups*32][N/2]  frame-major, code-bit order, two 4-bit two's-complement LLRs per byte
 *                (low nibble = even code bit)
 *   hard_packed  uint32[n_groups*32][N/32] bit n%32 of word n/32 = decoded bit n */
LDPC_B200_API int ldpc_b200_decode_packed(ldpc_b200_handle* h, const uint8_t* llr_packed, uint32_t* hard_packed, int n_groups,
                            int32_t* bf_iters, int32_t* its_per_group, int32_t* conv_iter);

/* float2LimitChar_4bit: q = clamp(trunc(x*scale), -7, 7). */
LDPC_B200_API int ldpc_b200_quantize(ldpc_b200_handle* h, const float* in, int8_t* out, int64_t length, float scale);
/* float2LimitChar_{1,2,3,4,5,6}bit (CLDPC.cpp:4385-4770): 6 = round to nearest even, clamp [-31,31]; 5 / 4 / 3 / 2 = truncate,
 * clamp [-16,15] / [-7,7] / [-4,3] / [-2,1]; 1 = +31 if trunc(x*scale) > 0 else -31. */
LDPC_B200_API int ldpc_b200_quantize_bits(ldpc_b200_handle* h, const float* in, int8_t* out, int64_t length, float scale, int bits);

/* Demap + de-interleave + regroup + quantise noisy symbols (complex64 interleaved re,im;
 * 32*N/modType symbols per group) into fixInput layout.  llr_float (optional) receives DeInterLeaveSeq. */
LDPC_B200_API int ldpc_b200_demap(ldpc_b200_handle* h, const float* symbols, int n_groups, float* llr_float, int8_t* fixInput);

/* Fused producer: interleave + map outputBits (NULL = all-zero codeword... see INTEGRATION.md), add
 * Philox4x32-10 AWGN for Eb/N0 = ebn0_db (sigma as CSimulate::Configure, CSimulate.cpp:67-75), demap,
 * de-interleave, quantise.  Frame i of the call uses Philox subsequence first_frame_index + i, so the
 * stream is independent of the GPU count.  symbols_out (optional) receives the noisy symbols. */
LDPC_B200_API int ldpc_b200_generate(ldpc_b200_handle* h, const int8_t* outputBits, float ebn0_db, uint64_t seed,
                       uint64_t first_frame_index, int n_groups, float* symbols_out, int8_t* fixInput);

/* CLDPC::GenMsgSeq (CLDPC.cpp:60-66; rand()%2 -> Philox4x32-10): info bits int8[32*K] per group.  Frame i of the call
 * uses Philox subsequence first_frame_index + i -- the same bits ldpc_b200_simulate draws for that frame index, so a
 * round can be replayed step by step (host/ldpc_sim.cpp does this to dump error frames). */
LDPC_B200_API int ldpc_b200_gen_msg_seq(ldpc_b200_handle* h, uint64_t seed, uint64_t first_frame_index, int n_groups, int8_t* inputBits);

/* Systematic encoder derived from H (the reference's GenMatrix is empty): inputBits int8[32*K] per group
 * -> outputBits int8[32*N] per group (two-region layout). */
LDPC_B200_API int ldpc_b200_encode(ldpc_b200_handle* h, const int8_t* inputBits, int8_t* outputBits, int n_groups);

/* CalculateErrors over n_groups groups: adds into counters[LDPC_B200_NUM_COUNTERS] (host pointer):
 * TEST_FRAME += 32*n_groups, ERROR_FRAME, ERROR_BITS (info bits only), LT3_ERR_BIT_FRAME. */
LDPC_B200_API int ldpc_b200_count_errors(ldpc_b200_handle* h, const int8_t* inputBits, const int8_t* decodedBits, int n_groups,
                           uint64_t* counters);

/* One Monte-Carlo round entirely on the device (CSimulate::Run): for n_groups groups draw info bits
 * (or use the fixed codeword when codeword != NULL, int8[N] = FakeEncoder), encode, map, add noise,
 * demap, quantise, decode, count.  Only the counters leave the GPU. */
LDPC_B200_API int ldpc_b200_simulate(ldpc_b200_handle* h, const int8_t* codeword, float ebn0_db, uint64_t seed,
                       uint64_t first_frame_index, int n_groups, uint64_t* counters);

/* Sum counters over ranks with one ncclAllReduce (main.cpp:170-182).  unique_id: 128 bytes from
 * ldpc_b200_nccl_unique_id on rank 0, distributed by the caller. */
LDPC_B200_API int ldpc_b200_nccl_unique_id(uint8_t unique_id[128]);
LDPC_B200_API int ldpc_b200_comm_init(ldpc_b200_handle* h, const uint8_t unique_id[128], int rank, int n_ranks);
LDPC_B200_API int ldpc_b200_allreduce_counters(ldpc_b200_handle* h, uint64_t* counters);

/* Pinned host memory for callers that want zero-copy-speed staging. */
LDPC_B200_API int ldpc_b200_host_alloc(void** ptr, uint64_t bytes);
LDPC_B200_API int ldpc_b200_host_free(void* ptr);

/* Timing of the last decode call, from CUDA events on the launching stream(s):
 * kernel_ms = sum of decoder-kernel durations, launches = number of kernels launched. */
LDPC_B200_API int ldpc_b200_last_timing(ldpc_b200_handle* h, float* kernel_ms, int32_t* launches);
/* Split of kernel_ms: message-passing kernel (decode_pair_kernel) vs group finalisation (finalize_kernel).  With more
 * than one chunk in flight the per-chunk durations overlap; create the handle with n_streams = 1 and
 * chunk_groups >= n_groups to time a kernel alone. */
LDPC_B200_API int ldpc_b200_last_timing_detail(ldpc_b200_handle* h, float* decode_ms, float* finalize_ms);
/* Host staging of ldpc_b200_decode() with HOST buffers (csrc/host_pack.h): `decodedBits` (CLDPC.h:124, one int8 per bit)
 * crosses PCIe as bits and is expanded into the caller's array by `threads` host threads while later chunks decode
 * (stage_out); optionally the int8 `fixInput` (CLDPC.h:123) crosses as nibbles (stage_in; chunks holding a value outside
 * [-8,7] go as bytes).  Pageable caller buffers are fine in staged directions.  Set when the handle is created:
 * LDPC_B200_HOST_THREADS (0 = off; default: this rank's share of the hardware threads, at most the CPUs of the GPU's NUMA node,
 * at most 64), LDPC_B200_STAGE_OUT (default 1 when the rank has >= 4 host threads) / LDPC_B200_STAGE_IN (default 1 when the
 * process is the only rank on the host (LOCAL_WORLD_SIZE) and has >= 8 cores -- with several links busy the box is bound by host
 * memory traffic, and packing needs twice what the copy as it is does).
 * LDPC_B200_HOST_REGISTER=1: pageable caller arrays are page-locked (cudaHostRegister) on first use and remembered by address
 * until destroy(), so that a caller that decodes out of one fixInput / decodedBits pair for the whole run (the reference does,
 * CLDPC.h:123-124) gets copy-engine transfers without changing its allocation; the arrays must then outlive the handle.
 * last_*_bytes: bytes the last ldpc_b200_decode / _decode_packed call moved over PCIe in each direction. */
LDPC_B200_API int ldpc_b200_host_staging(ldpc_b200_handle* h, int32_t* threads, int32_t* stage_in, int32_t* stage_out,
                                         uint64_t* last_h2d_bytes, uint64_t* last_d2h_bytes);

/* Bounds-check record (diagnostic).  A library built with -DLDPC_DEBUG_BOUNDS=1 (python build.py --out=... -DLDPC_DEBUG_BOUNDS=1)
 * range-checks, on the device, every shared-memory access of the message-passing code and every index into the LLR /
 * snapshot / hard-decision buffers; this returns the number of violations since create() and the first one
 * ((code << 32) | value).  compiled_in = 0 for the shipped library, whose kernels carry no checks (violations stays 0). */
LDPC_B200_API int ldpc_b200_debug_bounds(ldpc_b200_handle* h, int32_t* compiled_in, uint64_t* violations, uint64_t* first);

/* Hybrid host-buffer path (diagnostic): how many chunks of the last ldpc_b200_decode() call went through the host staging and
 * how many had their LLRs copied as they are by the copy engine.  With staging on and a PINNED fixInput both routes run at once --
 * the staged one is bound by the host's memory traffic, the direct one by the PCIe link: whenever a "direct" slot is idle the next
 * chunk takes it, otherwise the host threads pack it.  Decisions return as bits on both routes and are expanded by the host
 * threads (LDPC_B200_HYBRID_OUT_BITS=0: the direct slot copies decodedBits as bytes instead; needs that array pinned too).
 * LDPC_B200_HYBRID = number of direct slots (default 1, 0 = all staged). */
LDPC_B200_API int ldpc_b200_last_routing(ldpc_b200_handle* h, int32_t* staged_chunks, int32_t* direct_chunks);

/* NUMA placement chosen for the handle (diagnostic): node of the handle's GPU (-1 = unknown or disabled with LDPC_B200_NUMA=0)
 * and the number of CPUs of that node this process may use.  The staging threads run there and every pinned buffer the library
 * allocates -- its own staging mirrors and ldpc_b200_host_alloc(), which uses the CURRENT CUDA device -- is placed there, so
 * that host<->device copies do not cross the inter-socket link.  No counterpart in the reference (its threads are pinned to
 * cores by index, CSimulate.cpp:255-266). */
LDPC_B200_API int ldpc_b200_host_placement(ldpc_b200_handle* h, int32_t* numa_node, int32_t* numa_cpus);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
