"""Host staging helpers of the C-ABI's host-buffer path (csrc/host_pack.cpp): nibble packing of the reference's int8
`fixInput` layout and expansion of bit-packed decisions into `decodedBits`, against the numpy helpers of the package."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
N, K, M = 17664, 14592, 3072


@pytest.fixture(scope="module")
def hp(tmp_path_factory):
    csrc = ROOT / "mod-interleaveavx_multithreads-faid_b200" / "csrc"
    so = tmp_path_factory.mktemp("hp") / "libhostpack.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", f"-I{csrc}", f"-I{ROOT / 'include'}",
                    str(ROOT / "tools" / "probe" / "host_pack_capi.cpp"), str(csrc / "host_pack.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.hp_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.hp_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return lib


@pytest.mark.parametrize("groups,threads", [(1, 1), (3, 4), (5, 16)])
def test_pack_matches_numpy_helper(hp, groups, threads):
    import ldpc_b200
    rng = np.random.default_rng(groups)
    fix = rng.integers(-8, 8, size=(groups, 32 * N), dtype=np.int8)
    out = np.zeros((groups * 32, N // 2), dtype=np.uint8)
    assert hp.hp_pack(fix.ctypes.data, out.ctypes.data, groups, threads) == 1
    assert (out == ldpc_b200.pack_llr(fix)).all()
    # a value that does not fit a nibble is reported (the caller then ships bytes)
    for bad in (8, -9, 31, -128, 127):
        fix2 = fix.copy()
        fix2[groups - 1, int(rng.integers(0, 32 * N))] = bad
        assert hp.hp_pack(fix2.ctypes.data, out.ctypes.data, groups, threads) == 0


@pytest.mark.parametrize("frames,threads,misalign", [(1, 1, 0), (40, 4, 0), (100, 16, 16), (3, 2, 32), (33, 8, 48), (5, 3, 8), (5, 3, 56), (4, 2, 4), (4, 2, 1)])
def test_unpack_matches_numpy(hp, frames, threads, misalign):
    rng = np.random.default_rng(frames)
    hard = rng.integers(0, 2**32, size=(frames, N // 32), dtype=np.uint32)
    buf = np.zeros(frames * N + 192, dtype=np.int8)
    # 64-byte aligned: streaming stores; 8-byte aligned (malloc / numpy give 16): masked head and tail + streaming lines; else plain stores
    off = (-buf.ctypes.data) % 64 + 64 + misalign
    out = buf[off: off + frames * N]
    hp.hp_unpack(hard.ctypes.data, out.ctypes.data, frames, threads)
    ref = np.unpackbits(hard.view(np.uint8).reshape(frames, -1), axis=1, bitorder="little")
    assert (out.reshape(frames, N) == ref).all()
    assert not buf[:off].any() and not buf[off + frames * N:].any()


@pytest.mark.parametrize("threads,sleep_us", [(2, 0), (8, 50), (16, 400)])
def test_pool_fused_pass_under_stress(hp, threads, sleep_us):
    """host_stage_both (pack of one chunk + expansion of another in ONE pass over the pool) repeated on one pool, with pauses
    that let the workers fall asleep: every round must produce exactly the reference output (lost wake-ups or a worker that
    skips a generation would show as a hang or as stale 0xEE / 0x55 filler)."""
    import ldpc_b200
    rng = np.random.default_rng(threads)
    groups, frames = 2, 48
    fix = rng.integers(-8, 8, size=(groups, 32 * N), dtype=np.int8)
    packed_ref = np.ascontiguousarray(ldpc_b200.pack_llr(fix))
    hard = rng.integers(0, 2**32, size=(frames, N // 32), dtype=np.uint32)
    dec_ref = np.ascontiguousarray(np.unpackbits(hard.view(np.uint8).reshape(frames, -1), axis=1, bitorder="little").astype(np.int8))
    hp.hp_stress.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5
    assert hp.hp_stress(fix.ctypes.data, packed_ref.ctypes.data, hard.ctypes.data, dec_ref.ctypes.data, groups, frames, threads, 120, sleep_us) == 0
