"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  Oracle  -- oracle/_build/libldpc_oracle.so : this repo's plain-C restatement (oracle/ldpc_oracle.c)
  Ref     -- oracle/_ref/libldpc_ref_*.so    : the reference's own TUs compiled unmodified (+ harness)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(ROOT / "mod-interleaveavx_multithreads-faid_b200"))
from ldpc_b200.abi import Config, K, M, N  # noqa: E402  (struct layout only)

GROUP_BYTES = 32 * N


def build(ref=True):
    """Compile the C restatement and, when /root/reference is present, the reference build."""
    subprocess.run(["make", "-C", str(HERE), "oracle"], check=True, capture_output=True)
    if ref and Path("/root/reference/CLDPC.cpp").exists():
        subprocess.run(["make", "-C", str(HERE), "ref", "-j8"], check=True, capture_output=True)
        # the drop-in demonstration binary links the product library (tests/test_gpu_dropin_reference_frontend.py)
        if (ROOT / "mod-interleaveavx_multithreads-faid_b200" / "lib" / "libldpc_b200.so").exists():
            subprocess.run(["make", "-C", str(HERE), "dropin", "-j8"], check=True, capture_output=True)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Info(C.Structure):
    _fields_ = [
        ("iters_executed", C.c_int32),
        ("bf_iters", C.c_int32),
        ("conv_iter", C.c_int32 * 32),
        ("errsum_n", C.c_int32),
        ("errsum_log", (C.c_uint8 * 32) * 64),
    ]


class Oracle:
    def __init__(self):
        path = HERE / "_build" / "libldpc_oracle.so"
        if not path.exists():
            build(ref=False)
        self.lib = L = C.CDLL(str(path))
        L.ldpc_oracle_default_config.argtypes = [C.POINTER(Config), C.c_int, C.c_int]
        L.ldpc_oracle_decode.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.POINTER(Info)]
        L.ldpc_oracle_quantize_4bit.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int64]
        L.ldpc_oracle_quantize_bits.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_int]
        L.ldpc_oracle_transpose.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ldpc_oracle_itranspose.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ldpc_oracle_modulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ldpc_oracle_demodulate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ldpc_oracle_awgn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
        L.ldpc_oracle_bpsk_modulate.argtypes = [C.c_void_p, C.c_void_p]
        L.ldpc_oracle_sigma.argtypes = [C.c_float, C.c_int, C.c_double]
        L.ldpc_oracle_sigma.restype = C.c_float
        L.ldpc_oracle_encode_frame.argtypes = [C.c_void_p, C.c_void_p]
        L.ldpc_oracle_syndrome_weight.argtypes = [C.c_void_p]
        L.ldpc_oracle_encode_group.argtypes = [C.c_void_p, C.c_void_p]
        L.ldpc_oracle_calc_errors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ldpc_oracle_bench_decode.argtypes = [C.POINTER(Config), C.c_int, C.c_double, C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.ldpc_oracle_bench_decode.restype = C.c_double

    def default_config(self, method, lut=-1):
        cfg = Config()
        self.lib.ldpc_oracle_default_config(C.byref(cfg), method, lut)
        return cfg

    def decode(self, cfg, fix):
        """fix: int8[n_groups, 32*N] -> (decoded int8[n_groups, 32*N], list of Info)"""
        fix = np.ascontiguousarray(fix, dtype=np.int8).reshape(-1, GROUP_BYTES)
        out = np.empty_like(fix)
        infos = []
        for g in range(fix.shape[0]):
            info = Info()
            rc = self.lib.ldpc_oracle_decode(C.byref(cfg), _ptr(fix[g]), _ptr(out[g]), C.byref(info))
            assert rc == 0
            infos.append(info)
        return out, infos

    def quantize(self, x, scale, bits=4):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape, dtype=np.int8)
        if bits == 4:
            self.lib.ldpc_oracle_quantize_4bit(_ptr(out), _ptr(x), scale, x.size)
        else:
            assert self.lib.ldpc_oracle_quantize_bits(_ptr(out), _ptr(x), scale, x.size, bits) == 0
        return out

    def modulate(self, output_bits, mod_type, interleave):
        ob = np.ascontiguousarray(output_bits, dtype=np.int8)
        sym = np.empty(2 * 32 * N // mod_type, dtype=np.float32)
        assert self.lib.ldpc_oracle_modulate(_ptr(ob), mod_type, interleave, _ptr(sym)) == 0
        return sym

    def bpsk_modulate(self, output_bits):
        ob = np.ascontiguousarray(output_bits, dtype=np.int8)
        sym = np.empty(32 * N, dtype=np.float32)
        self.lib.ldpc_oracle_bpsk_modulate(_ptr(ob), _ptr(sym))
        return sym

    def demodulate(self, symbols, mod_type, interleave):
        s = np.ascontiguousarray(symbols, dtype=np.float32)
        demod = np.empty(32 * N, dtype=np.float32)
        deint = np.empty(32 * N, dtype=np.float32)
        assert self.lib.ldpc_oracle_demodulate(_ptr(s), mod_type, interleave, _ptr(demod), _ptr(deint)) == 0
        return demod, deint

    def awgn(self, symbols, sigma_d, state):
        s = np.ascontiguousarray(symbols, dtype=np.float32)
        out = np.empty_like(s)
        st = np.array(state, dtype=np.uint64)
        self.lib.ldpc_oracle_awgn(_ptr(s), _ptr(out), s.size // 2, sigma_d, _ptr(st))
        return out, st

    def sigma(self, ebn0, mod_type, rate=0.8444444):
        return float(self.lib.ldpc_oracle_sigma(ebn0, mod_type, rate))

    def encode_frame(self, info):
        i = np.ascontiguousarray(info, dtype=np.int8)
        cw = np.empty(N, dtype=np.int8)
        self.lib.ldpc_oracle_encode_frame(_ptr(i), _ptr(cw))
        return cw

    def encode_group(self, input_bits):
        i = np.ascontiguousarray(input_bits, dtype=np.int8)
        out = np.empty(32 * N, dtype=np.int8)
        self.lib.ldpc_oracle_encode_group(_ptr(i), _ptr(out))
        return out

    def syndrome_weight(self, cw):
        c = np.ascontiguousarray(cw, dtype=np.int8)
        return int(self.lib.ldpc_oracle_syndrome_weight(_ptr(c)))

    def calc_errors(self, input_bits, decoded):
        a = np.ascontiguousarray(input_bits, dtype=np.int8)
        d = np.ascontiguousarray(decoded, dtype=np.int8)
        st = np.zeros(3, dtype=np.uint64)
        self.lib.ldpc_oracle_calc_errors(_ptr(a), _ptr(d), _ptr(st))
        return st

    def bench_decode(self, cfg, groups, n_threads, min_seconds):
        g = np.ascontiguousarray(groups, dtype=np.int8).reshape(-1, GROUP_BYTES)
        done = C.c_int64(0)
        fps = self.lib.ldpc_oracle_bench_decode(C.byref(cfg), n_threads, min_seconds, _ptr(g), g.shape[0], C.byref(done))
        return fps, done.value


def ref_available(variant="faid3"):
    return (HERE / "_ref" / f"libldpc_ref_{variant}.so").exists()


class Ref:
    """The reference's own code.  variant: faid3 (as shipped) | faid2 | faid32 | instr (iteration-count instrumentation).

    The reference re-reads ./Profile.txt inside every decode call, so the process chdir()s into a private
    scratch directory for the lifetime of this object's calls (restored after each call)."""

    def __init__(self, variant="faid3"):
        path = HERE / "_ref" / f"libldpc_ref_{variant}.so"
        if not path.exists():
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(str(path), mode=os.RTLD_LOCAL)
        self.dir = tempfile.mkdtemp(prefix="ldpc_ref_")
        vp = C.c_void_p
        L.ref_write_profile.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float]
        L.ref_ldpc_create.restype = vp
        L.ref_ldpc_create.argtypes = [C.c_int]
        L.ref_ldpc_destroy.argtypes = [vp]
        L.ref_ldpc_set_iterations.argtypes = [vp, C.c_int]
        L.ref_decode.argtypes = [vp, C.c_int, vp, vp]
        L.ref_last_iterations.argtypes = [vp, C.c_int]
        L.ref_quantize_4bit.argtypes = [vp, vp, vp, C.c_float, C.c_int]
        L.ref_quantize_bits.argtypes = [vp, vp, vp, C.c_float, C.c_int, C.c_int]
        L.ref_transpose.argtypes = [vp, vp, C.c_int]
        L.ref_itranspose.argtypes = [vp, vp, C.c_int]
        L.ref_vn_weight.argtypes = [vp, vp]
        L.ref_sim_create.restype = vp
        L.ref_sim_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_sim_destroy.argtypes = [vp]
        L.ref_sim_ldpc.restype = vp
        L.ref_sim_ldpc.argtypes = [vp]
        L.ref_sim_rate.restype = C.c_double
        L.ref_sim_rate.argtypes = [vp]
        L.ref_sim_set_codeword.argtypes = [vp, vp, vp, vp, vp]
        L.ref_sim_set_output_bits.argtypes = [vp, vp, vp]
        L.ref_sim_noise_block.argtypes = [vp, C.c_float, C.c_float, vp, vp, vp, vp]
        L.ref_sim_demap_block.argtypes = [vp, vp, C.c_float, vp, vp, vp]
        L.ref_sim_rng_state.argtypes = [vp, vp]
        L.ref_sim_bpsk_modulate.argtypes = [vp, vp, vp]
        L.ref_sim_bpsk_receive.argtypes = [vp, vp, C.c_float, vp]
        L.ref_sim_decode_and_count.argtypes = [vp, C.c_int, vp, vp]
        L.ref_calc_errors.argtypes = [vp, vp, vp, vp, vp]
        L.ref_bench_decode.restype = C.c_double
        L.ref_bench_decode.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, vp, C.c_int, C.POINTER(C.c_long)]
        self._ldpc = {}

    class _Cwd:
        def __init__(self, d):
            self.d = d

        def __enter__(self):
            self.old = os.getcwd()
            os.chdir(self.d)

        def __exit__(self, *a):
            os.chdir(self.old)

    def write_profile(self, cfg):
        rc = self.lib.ref_write_profile(self.dir.encode(), cfg.snr_start, cfg.snr_pass, cfg.snr_end, cfg.decode_method,
                                        cfg.max_iteration, cfg.mod_type, cfg.interleave_mod_type, cfg.factor_1,
                                        cfg.factor_2, cfg.scale)
        assert rc == 0

    def _get_ldpc(self, max_iter):
        if max_iter not in self._ldpc:
            with self._Cwd(self.dir):
                self._ldpc[max_iter] = self.lib.ref_ldpc_create(max_iter)
        return self._ldpc[max_iter]

    def decode(self, cfg, fix, want_iters=False, method_override=None):
        """-> decoded int8[n_groups, 32*N], bf_iters list[, iterations list, errsum logs] (instr variant).
        method_override=101 calls CLDPC::Decode1 (dead code in the reference, generic puncture/shorten init)."""
        fix = np.ascontiguousarray(fix, dtype=np.int8).reshape(-1, GROUP_BYTES)
        out = np.empty_like(fix)
        self.write_profile(cfg)
        h = self._get_ldpc(cfg.max_iteration)
        bfs, its, logs = [], [], []
        with self._Cwd(self.dir):
            for g in range(fix.shape[0]):
                bfs.append(self.lib.ref_decode(h, cfg.decode_method if method_override is None else method_override, _ptr(fix[g]), _ptr(out[g])))
                if want_iters:
                    log = np.zeros((64, 32), dtype=np.uint8)
                    its.append(self.lib.ref_last_iterations(_ptr(log), 64))
                    logs.append(log)
        if want_iters:
            return out, bfs, its, logs
        return out, bfs

    def quantize(self, x, scale, bits=4):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape, dtype=np.int8)
        assert self.lib.ref_quantize_bits(self._get_ldpc(6), _ptr(out), _ptr(x), scale, x.size, bits) == 0
        return out

    def vn_weight(self):
        w = np.empty(N, dtype=np.int8)
        self.lib.ref_vn_weight(self._get_ldpc(6), _ptr(w))
        return w

    def bench_decode(self, cfg, groups, n_threads, min_seconds):
        g = np.ascontiguousarray(groups, dtype=np.int8).reshape(-1, GROUP_BYTES)
        self.write_profile(cfg)
        done = C.c_long(0)
        with self._Cwd(self.dir):
            fps = self.lib.ref_bench_decode(cfg.decode_method, cfg.max_iteration, n_threads, min_seconds, _ptr(g),
                                            g.shape[0], C.byref(done))
        return fps, done.value


class RefSim:
    """Mirror of one CSimulate object (CSimulate.cpp:41-180) on top of Ref."""

    def __init__(self, ref, cfg, seed=101):
        self.ref, self.cfg = ref, cfg
        ref.write_profile(cfg)
        with ref._Cwd(ref.dir):
            self.h = ref.lib.ref_sim_create(cfg.max_iteration, cfg.mod_type, cfg.interleave_mod_type, seed)
        self.nsym = 32 * N // cfg.mod_type

    def bpsk_modulate(self, output_bits):
        ob = np.ascontiguousarray(output_bits, dtype=np.int8)
        sym = np.empty(32 * N, dtype=np.float32)
        self.ref.lib.ref_sim_bpsk_modulate(self.h, _ptr(ob), _ptr(sym))
        return sym

    def bpsk_receive(self, noisy, scale):
        x = np.ascontiguousarray(noisy, dtype=np.float32)
        fix = np.empty(32 * N, dtype=np.int8)
        self.ref.lib.ref_sim_bpsk_receive(self.h, _ptr(x), scale, _ptr(fix))
        return fix

    def set_codeword(self, cw):
        cw = np.ascontiguousarray(cw, dtype=np.int8)
        inp = np.empty(32 * K, dtype=np.int8)
        outb = np.empty(32 * N, dtype=np.int8)
        mod = np.empty(2 * self.nsym, dtype=np.float32)
        self.ref.lib.ref_sim_set_codeword(self.h, _ptr(cw), _ptr(inp), _ptr(outb), _ptr(mod))
        return inp, outb, mod

    def set_output_bits(self, input_bits, output_bits):
        a = np.ascontiguousarray(input_bits, dtype=np.int8)
        b = np.ascontiguousarray(output_bits, dtype=np.int8)
        self.ref.lib.ref_sim_set_output_bits(self.h, _ptr(a), _ptr(b))

    def noise_block(self, sigma, scale):
        sym = np.empty(2 * self.nsym, dtype=np.float32)
        demod = np.empty(32 * N, dtype=np.float32)
        deint = np.empty(32 * N, dtype=np.float32)
        fix = np.empty(32 * N, dtype=np.int8)
        self.ref.lib.ref_sim_noise_block(self.h, sigma, scale, _ptr(sym), _ptr(demod), _ptr(deint), _ptr(fix))
        return sym, demod, deint, fix

    def demap_block(self, symbols, scale):
        s = np.ascontiguousarray(symbols, dtype=np.float32)
        demod = np.empty(32 * N, dtype=np.float32)
        deint = np.empty(32 * N, dtype=np.float32)
        fix = np.empty(32 * N, dtype=np.int8)
        self.ref.lib.ref_sim_demap_block(self.h, _ptr(s), scale, _ptr(demod), _ptr(deint), _ptr(fix))
        return demod, deint, fix

    def rng_state(self):
        st = np.zeros(3, dtype=np.uint64)
        self.ref.lib.ref_sim_rng_state(self.h, _ptr(st))
        return st

    def decode_and_count(self, method):
        dec = np.empty(32 * N, dtype=np.int8)
        st = np.zeros(3, dtype=np.uint64)
        with self.ref._Cwd(self.ref.dir):
            bf = self.ref.lib.ref_sim_decode_and_count(self.h, method, _ptr(dec), _ptr(st))
        return dec, st, bf
