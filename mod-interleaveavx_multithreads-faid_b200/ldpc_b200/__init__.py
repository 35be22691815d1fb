"""Host-side Python mirror of the reference's CLDPC / CSimulate surface over libldpc_b200.so (ctypes).

The shared library is the product; this module only marshals pointers.  It never computes on the CPU and
never imports anything from oracle/ -- if the CUDA library is missing or no B200 is visible, it raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import abi
from .abi import (BF_2B1C, BF_DTBF, BF_NONE, BF_PLAIN, FAID_2B1C, FAID_DTBF, GROUP, K, LUT_FAID2, LUT_FAID3,  # noqa: F401
                  LUT_FAID32, LUT_HYBRID, M, N, NMS, NUM_COUNTERS, OMS, OMS_BF, OMS_DTBF, Config)

PKG_DIR = Path(__file__).resolve().parent.parent
LIB_PATH = PKG_DIR / "lib" / "libldpc_b200.so"

_lib = None


class LdpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ldpc_b200 error {code}: {msg}")
        self.code = code


def load_library(path=None):
    """dlopen libldpc_b200.so and declare every exported symbol.  Fails loudly if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    if path is None and os.environ.get("LDPC_B200_LIB"):  # A/B experiments with alternative builds
        path = os.environ["LDPC_B200_LIB"]
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(f"{p} not found: build it with `python {PKG_DIR / 'build.py'}` (there is no CPU fallback)")
    lib = C.CDLL(str(p), mode=os.RTLD_LOCAL)
    for name, (res, args) in abi.EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise LdpcError(rc, load_library().ldpc_b200_last_error().decode())


def default_config(method, lut=-1):
    cfg = Config()
    _check(load_library().ldpc_b200_default_config(C.byref(cfg), method, lut))
    return cfg


def read_profile(path, cfg=None, lut=-1):
    cfg = cfg if cfg is not None else Config()
    _check(load_library().ldpc_b200_read_profile(str(path).encode(), C.byref(cfg), lut))
    return cfg


def _addr(x):
    """Raw address of a numpy array (host) or torch tensor (host or device)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    raise TypeError(type(x))


class PinnedArray:
    """numpy view over cudaMallocHost memory (ldpc_b200_host_alloc)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        _check(self.lib.ldpc_b200_host_alloc(C.byref(p), max(n, 1)))
        self.ptr = p
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.ldpc_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Decoder:
    """One engine handle = one CLDPC object of the reference (CLDPC.h:110-171): not thread-safe, one per GPU."""

    def __init__(self, cfg):
        self.lib = load_library()
        self.cfg = cfg
        h = C.c_void_p()
        _check(self.lib.ldpc_b200_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ldpc_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_factors(self, f1, f2):
        _check(self.lib.ldpc_b200_set_factors(self.h, f1, f2))

    # -- CLDPC::Decode* ---------------------------------------------------------------------------------
    def decode(self, fix, out=None, want_info=False):
        """fix: int8 [n_groups, 32*N] (reference fixInput layout; numpy or torch, host or device).
        Returns decodedBits int8 [n_groups, 32*N] (+ dict(bf_iters, its_per_group, conv_iter) if want_info)."""
        n_groups = int(np.prod(fix.shape)) // (32 * N)
        if out is None:
            if isinstance(fix, np.ndarray):
                out = np.empty((n_groups, 32 * N), dtype=np.int8)
            else:
                import torch
                out = torch.empty((n_groups, 32 * N), dtype=torch.int8, device=fix.device)
        info = None
        bf = its = conv = None
        if want_info:
            bf = np.zeros(n_groups, dtype=np.int32)
            its = np.zeros(n_groups, dtype=np.int32)
            conv = np.zeros(n_groups * 32, dtype=np.int32)
            info = dict(bf_iters=bf, its_per_group=its, conv_iter=conv.reshape(n_groups, 32))
        _check(self.lib.ldpc_b200_decode(self.h, _addr(fix), _addr(out), n_groups, _addr(bf), _addr(its), _addr(conv)))
        return (out, info) if want_info else out

    def decode_packed(self, llr_packed, out=None, want_info=False):
        """llr_packed: uint8 [frames, N/2]; returns uint32 [frames, N/32] packed hard decisions."""
        n_groups = int(np.prod(llr_packed.shape)) // (32 * N // 2)
        if out is None:
            if isinstance(llr_packed, np.ndarray):
                out = np.empty((n_groups * 32, N // 32), dtype=np.uint32)
            else:
                import torch
                out = torch.empty((n_groups * 32, N // 32), dtype=torch.int32, device=llr_packed.device)
        bf = its = conv = None
        info = None
        if want_info:
            bf = np.zeros(n_groups, dtype=np.int32)
            its = np.zeros(n_groups, dtype=np.int32)
            conv = np.zeros(n_groups * 32, dtype=np.int32)
            info = dict(bf_iters=bf, its_per_group=its, conv_iter=conv.reshape(n_groups, 32))
        _check(self.lib.ldpc_b200_decode_packed(self.h, _addr(llr_packed), _addr(out), n_groups, _addr(bf), _addr(its), _addr(conv)))
        return (out, info) if want_info else out

    def last_timing(self):
        ms = C.c_float(0)
        n = C.c_int32(0)
        _check(self.lib.ldpc_b200_last_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def host_staging(self):
        """-> dict(threads, stage_in, stage_out, last_h2d_bytes, last_d2h_bytes) of the host-buffer path (ldpc_b200_host_staging)"""
        t, i, o = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        a, b = C.c_uint64(0), C.c_uint64(0)
        _check(self.lib.ldpc_b200_host_staging(self.h, C.byref(t), C.byref(i), C.byref(o), C.byref(a), C.byref(b)))
        return {"threads": t.value, "stage_in": bool(i.value), "stage_out": bool(o.value), "last_h2d_bytes": a.value, "last_d2h_bytes": b.value}

    def debug_bounds(self):
        """-> dict(compiled_in, violations, first): device-side bounds-check record (ldpc_b200_debug_bounds)"""
        a, b, c = C.c_int32(0), C.c_uint64(0), C.c_uint64(0)
        _check(self.lib.ldpc_b200_debug_bounds(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"compiled_in": bool(a.value), "violations": b.value, "first": c.value}

    def last_routing(self):
        """-> dict(staged_chunks, direct_chunks) of the last decode() call with host buffers (ldpc_b200_last_routing)"""
        a, b = C.c_int32(0), C.c_int32(0)
        _check(self.lib.ldpc_b200_last_routing(self.h, C.byref(a), C.byref(b)))
        return {"staged_chunks": a.value, "direct_chunks": b.value}

    def host_placement(self):
        """-> dict(numa_node, numa_cpus): NUMA node of the handle's GPU and this process's CPUs on it (ldpc_b200_host_placement)"""
        a, b = C.c_int32(-1), C.c_int32(0)
        _check(self.lib.ldpc_b200_host_placement(self.h, C.byref(a), C.byref(b)))
        return {"numa_node": a.value, "numa_cpus": b.value}

    def last_timing_detail(self):
        a = C.c_float(0)
        b = C.c_float(0)
        _check(self.lib.ldpc_b200_last_timing_detail(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def pack_llr(fix):
    """reference fixInput layout int8 [n_groups, 32*N] -> native nibble layout uint8 [n_groups*32, N/2] (host helper)."""
    fix = np.asarray(fix, dtype=np.int8).reshape(-1, 32 * N)
    g = fix.shape[0]
    info = fix[:, : 32 * K].reshape(g, 32, K)
    par = fix[:, 32 * K:].reshape(g, 32, M)
    frames = np.concatenate([info, par], axis=2).reshape(g * 32, N)
    lo = frames[:, 0::2].astype(np.uint8) & 0xF
    hi = frames[:, 1::2].astype(np.uint8) & 0xF
    return (lo | (hi << 4)).astype(np.uint8)


def unpack_hard(hard_packed):
    """uint32 [frames, N/32] -> int8 [frames, N] of 0/1 (host helper)."""
    hp = np.ascontiguousarray(hard_packed).view(np.uint32).reshape(-1, N // 32)
    bits = np.unpackbits(hp.view(np.uint8), bitorder="little").reshape(hp.shape[0], N)
    return bits.astype(np.int8)


def _new_like(x, shape, np_dtype, torch_dtype_name):
    if isinstance(x, np.ndarray) or x is None:
        return np.empty(shape, dtype=np_dtype)
    import torch
    return torch.empty(shape, dtype=getattr(torch, torch_dtype_name), device=x.device)


def _decoder_methods():
    """Frame-generation / scoring entry points (CModulate, CChannel, CLDPC::Encode, CalculateErrors, CSimulate::Run)."""

    def quantize(self, x, scale=None, out=None, bits=4):
        """CLDPC::float2LimitChar_{bits}bit (4 = the one CSimulate::Run calls)"""
        scale = self.cfg.scale if scale is None else scale
        n = int(np.prod(x.shape))
        out = _new_like(x, x.shape, np.int8, "int8") if out is None else out
        _check(self.lib.ldpc_b200_quantize_bits(self.h, _addr(x), _addr(out), n, scale, bits))
        return out

    def demap(self, symbols, want_float=True):
        """CModulate::Demodulation + AfterDeModulationDeInterleaver + float2LimitChar_4bit.
        symbols: float32 [n_groups, 2*32*N/modType] (re,im interleaved)."""
        per_group = 2 * 32 * N // self.cfg.mod_type
        n_groups = int(np.prod(symbols.shape)) // per_group
        llr = _new_like(symbols, (n_groups, 32 * N), np.float32, "float32") if want_float else None
        fix = _new_like(symbols, (n_groups, 32 * N), np.int8, "int8")
        _check(self.lib.ldpc_b200_demap(self.h, _addr(symbols), n_groups, _addr(llr), _addr(fix)))
        return llr, fix

    def generate(self, output_bits, ebn0_db, seed, first_frame, n_groups, want_symbols=False, like=None):
        """Fused producer: map + Philox AWGN + demap + de-interleave + quantise -> fixInput int8 [n_groups, 32*N]."""
        ref = output_bits if output_bits is not None else like
        fix = _new_like(ref, (n_groups, 32 * N), np.int8, "int8")
        sym = None
        if want_symbols:
            per_group = 32 * N if self.cfg.mod_type == 1 else 2 * 32 * N // self.cfg.mod_type
            sym = _new_like(ref, (n_groups, per_group), np.float32, "float32")
        _check(self.lib.ldpc_b200_generate(self.h, _addr(output_bits), ebn0_db, seed, first_frame, n_groups, _addr(sym), _addr(fix)))
        return (fix, sym) if want_symbols else fix

    def gen_msg_seq(self, seed, first_frame, n_groups):
        """CLDPC::GenMsgSeq: int8 [n_groups, 32*K] info bits (the ones simulate() draws for these frame indices)."""
        out = np.empty((n_groups, 32 * K), dtype=np.int8)
        _check(self.lib.ldpc_b200_gen_msg_seq(self.h, seed, first_frame, n_groups, _addr(out)))
        return out

    def encode(self, input_bits):
        """CLDPC::Encode: int8 [n_groups, 32*K] -> int8 [n_groups, 32*N] (two-region layout)."""
        n_groups = int(np.prod(input_bits.shape)) // (32 * K)
        out = _new_like(input_bits, (n_groups, 32 * N), np.int8, "int8")
        _check(self.lib.ldpc_b200_encode(self.h, _addr(input_bits), _addr(out), n_groups))
        return out

    def count_errors(self, input_bits, decoded, counters=None):
        """CLDPC::CalculateErrors; returns the uint64[NUM_COUNTERS] vector (accumulated into `counters` if given)."""
        n_groups = int(np.prod(decoded.shape)) // (32 * N)
        counters = np.zeros(NUM_COUNTERS, dtype=np.uint64) if counters is None else counters
        _check(self.lib.ldpc_b200_count_errors(self.h, _addr(input_bits), _addr(decoded), n_groups, _addr(counters)))
        return counters

    def simulate(self, ebn0_db, seed, first_frame, n_groups, codeword=None, counters=None):
        """One Monte-Carlo round on the device (CSimulate::Run); only the counters come back."""
        counters = np.zeros(NUM_COUNTERS, dtype=np.uint64) if counters is None else counters
        cw = None if codeword is None else np.ascontiguousarray(codeword, dtype=np.int8)
        _check(self.lib.ldpc_b200_simulate(self.h, _addr(cw), ebn0_db, seed, first_frame, n_groups, _addr(counters)))
        return counters

    def comm_init(self, unique_id, rank, n_ranks):
        uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        _check(self.lib.ldpc_b200_comm_init(self.h, _addr(uid), rank, n_ranks))

    def allreduce_counters(self, counters):
        _check(self.lib.ldpc_b200_allreduce_counters(self.h, _addr(counters)))
        return counters

    for f in (quantize, demap, generate, gen_msg_seq, encode, count_errors, simulate, comm_init, allreduce_counters):
        setattr(Decoder, f.__name__, f)


_decoder_methods()


def nccl_unique_id():
    uid = np.zeros(128, dtype=np.uint8)
    _check(load_library().ldpc_b200_nccl_unique_id(_addr(uid)))
    return bytes(uid)


def ebn0_sigma(cfg, ebn0_db):
    """sigma of CSimulate::Configure (CSimulate.cpp:67-75)."""
    m = cfg.mod_type
    if m == 1:
        return np.float32(1.0 / np.sqrt(2.0 * cfg.code_rate * m * 10.0 ** (0.1 * ebn0_db)))
    return np.float32(1.0 / np.sqrt(cfg.code_rate * m * 10.0 ** (0.1 * ebn0_db)))
