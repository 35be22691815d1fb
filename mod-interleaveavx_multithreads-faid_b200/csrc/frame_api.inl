// frame_api.inl -- C-ABI entry points for frame generation, encoding, counting (included by ldpc_b200.cu)
extern "C" {
#define LDPC_NOT_YET(name) return fail(LDPC_B200_EINVAL, name ": not implemented in this build")
int ldpc_b200_quantize(ldpc_b200_handle*, const float*, int8_t*, int64_t, float) { LDPC_NOT_YET("quantize"); }
int ldpc_b200_demap(ldpc_b200_handle*, const float*, int, float*, int8_t*) { LDPC_NOT_YET("demap"); }
int ldpc_b200_generate(ldpc_b200_handle*, const int8_t*, float, uint64_t, uint64_t, int, float*, int8_t*) { LDPC_NOT_YET("generate"); }
int ldpc_b200_encode(ldpc_b200_handle*, const int8_t*, int8_t*, int) { LDPC_NOT_YET("encode"); }
int ldpc_b200_count_errors(ldpc_b200_handle*, const int8_t*, const int8_t*, int, uint64_t*) { LDPC_NOT_YET("count_errors"); }
int ldpc_b200_simulate(ldpc_b200_handle*, const int8_t*, float, uint64_t, uint64_t, int, uint64_t*) { LDPC_NOT_YET("simulate"); }
int ldpc_b200_nccl_unique_id(uint8_t*) { LDPC_NOT_YET("nccl_unique_id"); }
int ldpc_b200_comm_init(ldpc_b200_handle*, const uint8_t*, int, int) { LDPC_NOT_YET("comm_init"); }
int ldpc_b200_allreduce_counters(ldpc_b200_handle*, uint64_t*) { LDPC_NOT_YET("allreduce_counters"); }
}
