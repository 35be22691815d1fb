"""ctypes mirror of include/ldpc_b200.h (the C-ABI of libldpc_b200.so).

Only declarations live here: struct layout, constants, and the argtypes of every exported symbol.
`EXPORTS` is what tests/test_abi.py checks the shared library against.
"""
import ctypes as C

N = 17664
M = 3072
K = 14592
GROUP = 32
ABI_VERSION = 1
NUM_COUNTERS = 128
CNT_TEST_FRAME, CNT_ERROR_FRAME, CNT_ERROR_BITS, CNT_LT3, CNT_GROUPS, CNT_MS_ITERS_SUM = 0, 1, 2, 3, 4, 5
CNT_BF_HIST = 8
CNT_MS_HIST = 64

NMS, OMS, FAID_DTBF, OMS_BF, OMS_DTBF, FAID_2B1C = range(6)
LUT_FAID3, LUT_FAID32, LUT_FAID2, LUT_HYBRID = range(4)
BF_NONE, BF_PLAIN, BF_DTBF, BF_2B1C = range(4)

OK, EINVAL, ECUDA, ENOMEM, EIO, ENODEV, ENCCL = 0, -1, -2, -3, -4, -5, -6


class Config(C.Structure):
    """ldpc_b200_config (include/ldpc_b200.h): Profile.txt fields + the reference's compile-time constants."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("abi_version", C.c_uint32),
        ("snr_start", C.c_float),
        ("snr_pass", C.c_float),
        ("snr_end", C.c_float),
        ("decode_method", C.c_int32),
        ("max_iteration", C.c_int32),
        ("mod_type", C.c_int32),
        ("interleave_mod_type", C.c_int32),
        ("factor_1", C.c_int32),
        ("factor_2", C.c_int32),
        ("nb_frames", C.c_int32),
        ("scale", C.c_float),
        ("Z", C.c_int32),
        ("v2c_lut", C.c_int8 * 8 * 4 * 6),
        ("v2c_lut_ef", C.c_int8 * 8 * 4 * 6),
        ("ef_elimination", C.c_int32),
        ("ef_floor_err_count", C.c_int32),
        ("ef_floor_iter_thresh", C.c_int32),
        ("oms_floor_err_count", C.c_int32),
        ("oms_floor_iter_thresh", C.c_int32),
        ("bf_mode", C.c_int32),
        ("bf_max_iter", C.c_int32),
        ("dtbf_L0", C.c_int32),
        ("dtbf_L1", C.c_int32),
        ("dtbf_delta", C.c_int32),
        ("dtbf_alpha", C.c_int32),
        ("regular_col_weight", C.c_int32),
        ("hard2_threshold", C.c_int32),
        ("puncture_tail", C.c_int32),
        ("code_rate", C.c_double),
        ("device", C.c_int32),
        ("n_streams", C.c_int32),
        ("chunk_groups", C.c_int32),
        ("quant_bits", C.c_int32),
        ("oms_mode", C.c_int32),
        ("oms_offset", C.c_int32),
        ("codeword_reuse", C.c_int32),
        ("reserved", C.c_int32 * 1),
    ]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            if name in ("v2c_lut", "v2c_lut_ef"):
                v = [[list(r) for r in it] for it in v]
            elif name == "reserved":
                v = list(v)
            out[name] = v
        return out


_p = C.c_void_p
_i8p = C.c_void_p  # raw addresses: host or device pointers
_cfgp = C.POINTER(Config)

# symbol -> (restype, argtypes)
EXPORTS = {
    "ldpc_b200_version": (C.c_char_p, []),
    "ldpc_b200_last_error": (C.c_char_p, []),
    "ldpc_b200_default_config": (C.c_int, [_cfgp, C.c_int, C.c_int]),
    "ldpc_b200_read_profile": (C.c_int, [C.c_char_p, _cfgp, C.c_int]),
    "ldpc_b200_create": (C.c_int, [_cfgp, C.POINTER(_p)]),
    "ldpc_b200_destroy": (C.c_int, [_p]),
    "ldpc_b200_set_factors": (C.c_int, [_p, C.c_int, C.c_int]),
    "ldpc_b200_set_max_iteration": (C.c_int, [_p, C.c_int]),
    "ldpc_b200_decode": (C.c_int, [_p, _i8p, _i8p, C.c_int, _p, _p, _p]),
    "ldpc_b200_decode_packed": (C.c_int, [_p, _i8p, _i8p, C.c_int, _p, _p, _p]),
    "ldpc_b200_quantize": (C.c_int, [_p, _p, _p, C.c_int64, C.c_float]),
    "ldpc_b200_quantize_bits": (C.c_int, [_p, _p, _p, C.c_int64, C.c_float, C.c_int]),
    "ldpc_b200_demap": (C.c_int, [_p, _p, C.c_int, _p, _p]),
    "ldpc_b200_generate": (C.c_int, [_p, _p, C.c_float, C.c_uint64, C.c_uint64, C.c_int, _p, _p]),
    "ldpc_b200_encode": (C.c_int, [_p, _p, _p, C.c_int]),
    "ldpc_b200_gen_msg_seq": (C.c_int, [_p, C.c_uint64, C.c_uint64, C.c_int, _p]),
    "ldpc_b200_count_errors": (C.c_int, [_p, _p, _p, C.c_int, _p]),
    "ldpc_b200_simulate": (C.c_int, [_p, _p, C.c_float, C.c_uint64, C.c_uint64, C.c_int, _p]),
    "ldpc_b200_nccl_unique_id": (C.c_int, [_p]),
    "ldpc_b200_comm_init": (C.c_int, [_p, _p, C.c_int, C.c_int]),
    "ldpc_b200_allreduce_counters": (C.c_int, [_p, _p]),
    "ldpc_b200_host_alloc": (C.c_int, [C.POINTER(_p), C.c_uint64]),
    "ldpc_b200_host_free": (C.c_int, [_p]),
    "ldpc_b200_last_timing": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "ldpc_b200_last_timing_detail": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ldpc_b200_debug_bounds": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ldpc_b200_last_routing": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ldpc_b200_host_placement": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ldpc_b200_host_staging": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}
