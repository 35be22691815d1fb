"""CPU check of the EXACT integer arithmetic of the CUDA message-passing kernel.

tools/emu/emu_decode.cpp compiles the per-layer device functions of csrc/decode_kernels.cuh unchanged for the host
(the sm_100a SIMD intrinsics are provided by tools/emu/cuda_emu_shim.h) and runs them thread after thread.  The
result -- hard decisions at the group's stop point, executed iterations, per-frame convergence iteration -- must
equal the oracle's with the bit-flipping stage switched off.  This is what lets a change of the kernel's bias /
packing tricks be validated in the GPU-less container; the `-m gpu` tests then confirm the real kernel.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import llrgen

ROOT = Path(__file__).resolve().parent.parent
EMU_DIR = ROOT / "tools" / "emu"
N = 17664


@pytest.fixture(scope="module")
def emu():
    src = EMU_DIR / "emu_decode.cpp"
    so = EMU_DIR / "libemu_decode.so"
    csrc = ROOT / "mod-interleaveavx_multithreads-faid_b200" / "csrc"
    deps = [src, EMU_DIR / "cuda_emu_shim.h", csrc / "decode_kernels.cuh", csrc / "host_params.h", ROOT / "include" / "ldpc_code_tables.h"]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", f"-I{EMU_DIR}", f"-I{ROOT / 'include'}", f"-I{csrc}",
                        "-o", str(so), str(src)], check=True)
    lib = C.CDLL(str(so))
    lib.emu_decode_group.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
    return lib


def run_emu(lib, cfg, fix_group, allow_fast=1):
    """-> hard decisions, executed iterations, per-frame convergence iteration, MONO flag, kernel kind that ran"""
    out = np.empty(32 * N, dtype=np.int8)
    its = C.c_int32(0)
    conv = np.empty(32, dtype=np.int32)
    rc = lib.emu_decode_group(C.byref(cfg), allow_fast, fix_group.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                              C.byref(its), conv.ctypes.data_as(C.c_void_p))
    return out, its.value, conv, rc & 1, rc >> 4


CASES = [
    # method, lut, factors (None = shipped), max_iter, Eb/N0 points
    (0, -1, None, 6, (3.0, 3.8)),
    (0, -1, (22, 29), 6, (3.4,)),
    (0, -1, (29, 22), 4, (3.4,)),      # cste_1 < cste_2 possible: exercises the mask-select (non-MONO) variant
    (1, -1, None, 6, (3.0, 4.4)),
    (1, -1, (2, 5), 10, (3.6,)),
    (1, -1, "oms_mode0", 6, (3.3, 3.9)),   # OMS_MODE 0: cste = min - 1, negative for min = 0
    (2, 0, None, 6, (3.0, 4.4)),
    (2, 1, None, 6, (3.5,)),
    (2, 2, None, 12, (3.7,)),
    (5, 3, None, 6, (3.0, 3.7, 4.4)),
]


@pytest.mark.parametrize("method,lut,factors,max_iter,ebs", CASES)
def test_kernel_arithmetic_matches_oracle(oracle, emu, method, lut, factors, max_iter, ebs):
    cfg = oracle.default_config(method, lut)
    cfg.max_iteration = max_iter
    if factors == "oms_mode0":
        cfg.oms_mode = 0
    elif factors:
        cfg.factor_1, cfg.factor_2 = factors
    cfg.bf_mode = 0  # compare the min-sum stage only: the BF stage is a separate kernel (bf_kernels.cuh)
    cfg.bf_max_iter = 0
    for i, eb in enumerate(ebs):
        fix, _ = llrgen.qpsk_llr_groups(1, eb, scale=cfg.scale, seed=1000 + 10 * method + i)
        ref, infos = oracle.decode(cfg, fix)
        # FAID methods: both the monotone-LUT fast path (kinds 4/5) and the general per-edge-LUT path (kinds 2/3)
        for allow_fast in ((1, 0) if method in (2, 5) else (1,)):
            out, its, conv, mono, kind = run_emu(emu, cfg, fix[0], allow_fast)
            if factors == (29, 22):
                assert mono == 0
            if method in (2, 5):
                assert kind == {(2, 1): 4, (2, 0): 2, (5, 1): 5, (5, 0): 3}[(method, allow_fast)]
            nd = int((out != ref[0]).sum())
            assert nd == 0, f"method {method} lut {lut} @ {eb} dB (kind {kind}): {nd} differing bits"
            if method != 0:
                assert its == infos[0].iters_executed
                assert list(conv) == list(infos[0].conv_iter)


def test_extreme_inputs(oracle, emu):
    """All-saturated and all-zero LLRs, alternating signs: the clamps of the biased representation."""
    rng = np.random.default_rng(5)
    for method in (0, 1, 2, 5):
        cfg = oracle.default_config(method, -1)
        cfg.bf_mode = 0
        cfg.bf_max_iter = 0
        for kind in range(3):
            if kind == 0:
                fix = rng.choice(np.array([-7, 7], dtype=np.int8), size=(1, 32 * N))
            elif kind == 1:
                fix = np.zeros((1, 32 * N), dtype=np.int8)
            else:
                fix = rng.integers(-7, 8, size=(1, 32 * N), dtype=np.int8)
            ref, _ = oracle.decode(cfg, fix)
            for allow_fast in ((1, 0) if method in (2, 5) else (1,)):
                out = run_emu(emu, cfg, fix[0], allow_fast)[0]
                assert int((out != ref[0]).sum()) == 0, (method, kind, allow_fast)


@pytest.mark.parametrize("mag_ok,mag_bad,n_flip,max_iter", [(2, 7, 4, 2), (1, 7, 5, 3), (1, 7, 6, 6)])
def test_erasure_mode_kernel_arithmetic(oracle, emu, mag_ok, mag_bad, n_flip, max_iter):
    """EF_ELIMINATION 2 (KIND_FAID_ER): the kernel's erasure of weight-3 variable nodes against the oracle, on the inputs
    of tests/test_oracle_vs_reference.py::test_erasure_mode_ef_elimination_2 (where the oracle is pinned to the reference)
    plus a noisy group; min-sum stage only."""
    cfg = oracle.default_config(2, 0)
    cfg.max_iteration = max_iter
    cfg.ef_elimination, cfg.ef_floor_err_count, cfg.ef_floor_iter_thresh = 2, 20, 6
    cfg.bf_mode = 0
    cfg.bf_max_iter = 0
    fixes = [llrgen.sparse_error_groups(mag_ok, mag_bad, n_flip, seed=1), llrgen.qpsk_llr_groups(1, 3.9, seed=77)[0]]
    changed = False
    for fix in fixes:
        ref, infos = oracle.decode(cfg, fix)
        out, its, conv, mono, kind = run_emu(emu, cfg, fix[0], 1)
        assert kind == 6
        assert int((out != ref[0]).sum()) == 0
        assert its == infos[0].iters_executed
        assert list(conv) == list(infos[0].conv_iter)
        cfg.ef_elimination = 1
        changed |= bool((oracle.decode(cfg, fix)[0] != ref).any())
        cfg.ef_elimination = 2
    assert changed, "the erasures must have fired"
    # the hybrid decoder has no erasure: EF_ELIMINATION 2 runs its error-floor kinds
    cfg5 = oracle.default_config(5, -1)
    cfg5.max_iteration = max_iter
    cfg5.ef_elimination, cfg5.ef_floor_err_count, cfg5.ef_floor_iter_thresh = 2, 20, 6
    cfg5.bf_mode = 0
    cfg5.bf_max_iter = 0
    ref, _ = oracle.decode(cfg5, fixes[0])
    for allow_fast in (1, 0):
        out, _, _, _, kind = run_emu(emu, cfg5, fixes[0][0], allow_fast)
        assert kind == (5 if allow_fast else 3)
        assert int((out != ref[0]).sum()) == 0


@pytest.mark.parametrize("method", [0, 1, 2, 5])
def test_random_configurations_kernel_arithmetic(oracle, emu, method):
    """The draws of tests/test_gpu_random_configs.py (random factors, LUTs, thresholds, iteration limits, puncturing) on the
    CPU emulation of the kernel arithmetic; min-sum stage only."""
    from test_gpu_random_configs import _randomise
    rng = np.random.default_rng(4000 + method)
    fix, _ = llrgen.qpsk_llr_groups(1, 3.5, seed=600 + method)
    for trial in range(2):
        cfg = oracle.default_config(method, -1)
        draw = _randomise((cfg,), method, rng)
        cfg.max_iteration = min(cfg.max_iteration, 6)
        cfg.bf_mode = 0
        cfg.bf_max_iter = 0
        ref, infos = oracle.decode(cfg, fix)
        out, its, conv, mono, kind = run_emu(emu, cfg, fix[0], 1)
        assert int((out != ref[0]).sum()) == 0, (kind, draw)
        if method != 0:
            assert its == infos[0].iters_executed
