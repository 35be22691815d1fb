"""Seeded synthetic decoder inputs shared by the tests, smoke() and bench.py (numpy only, no oracle needed).

QPSK over AWGN of a given codeword, demapped/quantised exactly as the reference chain does
(CSimulate.cpp:67-75,126-132; CModulate.cpp:270-280; CLDPC.cpp:4524-4582) but with numpy's PCG64 noise.
"""
from pathlib import Path

import numpy as np

N, M, K = 17664, 3072, 14592
ROOT = Path(__file__).resolve().parent.parent


def golden_codeword():
    return np.unpackbits(np.load(ROOT / "tests" / "golden" / "codeword_50gpon.npy"))[:N].astype(np.int8)


def sigma(ebn0_db, mod_type=2, rate=0.8444444):
    return np.float32(1.0 / np.sqrt(rate * mod_type * 10.0 ** (0.1 * ebn0_db)))


def qpsk_llr_groups(n_groups, ebn0_db, scale=13.0, seed=1, codeword=None):
    """-> fixInput int8 [n_groups, 32*N] in the reference two-region layout, tx bits int8 [N]."""
    cw = golden_codeword() if codeword is None else np.asarray(codeword, dtype=np.int8)
    rng = np.random.default_rng(seed)
    amp = np.float32(0.707107)
    tx = (2.0 * cw.astype(np.float32) - 1.0) * amp
    sd = np.float32(sigma(ebn0_db) / np.sqrt(2.0))
    out = np.empty((n_groups, 32 * N), dtype=np.int8)
    for g in range(n_groups):
        y = tx[None, :] + rng.standard_normal((32, N), dtype=np.float32) * sd
        q = np.clip(np.trunc(y * np.float32(scale)), -7, 7).astype(np.int8)
        out[g, : 32 * K] = q[:, :K].reshape(-1)
        out[g, 32 * K:] = q[:, K:].reshape(-1)
    return out, cw


def sparse_error_groups(mag_ok, mag_bad, n_flip, seed=1, codeword=None):
    """One group whose frames are the codeword at LLR magnitude `mag_ok` with `n_flip` bits of the weight-3 block columns
    (17..66) set to the WRONG sign at magnitude `mag_bad`: few unsatisfied checks, each flipped VN with all three of its
    checks unsatisfied -- the operating point of the reference's error-floor logic (EF_ELIMINATION 1 / 2).
    -> fixInput int8 [1, 32*N] (two-region layout)."""
    cw = golden_codeword() if codeword is None else np.asarray(codeword, dtype=np.int8)
    rng = np.random.default_rng(seed)
    frames = np.empty((32, N), dtype=np.int8)
    for f in range(32):
        llr = (2 * cw.astype(np.int16) - 1) * mag_ok
        idx = rng.choice(np.arange(17 * 256, 67 * 256), size=n_flip, replace=False)
        llr[idx] = -(2 * cw[idx].astype(np.int16) - 1) * mag_bad
        frames[f] = llr
    out = np.empty((1, 32 * N), dtype=np.int8)
    out[0, : 32 * K] = frames[:, :K].reshape(-1)
    out[0, 32 * K:] = frames[:, K:].reshape(-1)
    return out
