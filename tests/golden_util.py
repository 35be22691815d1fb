"""Loader for tests/golden/decode_vectors.npz (written by tools/make_golden.py from the reference build)."""
from pathlib import Path

import numpy as np

import llrgen

ROOT = Path(__file__).resolve().parent.parent
N, M, K = 17664, 3072, 14592
CASE_NAMES = ["nms", "nms_f22_29", "nms_it12", "oms", "oms_f2_5", "faid3_dtbf", "faid32_dtbf", "faid2_dtbf",
              "faid3_dtbf_it12", "oms_bf", "oms_dtbf", "hybrid_2b1c"]
_cache = {}


def vectors():
    if "v" not in _cache:
        _cache["v"] = np.load(ROOT / "tests" / "golden" / "decode_vectors.npz")
    return _cache["v"]


def unpack_nibbles(p):
    lo = ((p & 0xF).astype(np.int8) ^ 8) - 8
    hi = ((p >> 4).astype(np.int8) ^ 8) - 8
    out = np.empty(p.shape[:-1] + (p.shape[-1] * 2,), dtype=np.int8)
    out[..., 0::2] = lo
    out[..., 1::2] = hi
    return out


def case(name):
    """-> dict(method, lut, max_iteration, factor_1, factor_2, scale, fix [3, 32N], dec [3, 32N], bf, its, errsum)"""
    v = vectors()
    method, lut, mi, f1, f2 = [int(x) for x in v[f"{name}.cfg"]]
    scale = float(v[f"{name}.scale"][0])
    fix = unpack_nibbles(v[f"llr_scale{scale}"])
    cwg = np.tile(llrgen.golden_codeword(), 32)
    dec = np.stack([np.unpackbits(d)[: 32 * N].astype(np.int8) ^ cwg for d in v[f"{name}.dec_xor_cw"]])
    return dict(method=method, lut=lut, max_iteration=mi, factor_1=f1, factor_2=f2, scale=scale, fix=fix, dec=dec,
                bf=v[f"{name}.bf"], its=v[f"{name}.its"], errsum=v[f"{name}.errsum"])


def apply(cfg, c):
    cfg.max_iteration, cfg.factor_1, cfg.factor_2, cfg.scale = c["max_iteration"], c["factor_1"], c["factor_2"], c["scale"]
    return cfg
