"""Reference-side Monte-Carlo chain for the BER/FER confidence-interval gates (test infrastructure).

Runs the reference's own CModulate + CChannel (3-LCG + Box-Muller, CChannel.cpp:71-97) + CLDPC objects (oracle/_ref, compiled
unmodified) exactly like CSimulate::Run does, as `len(seeds)` independent "threads" with the per-thread seeds of
CSimulate.cpp:11 (101, 103, 107, ...), spread over a process pool.  The result depends on (seeds, blocks_per_seed) only, not on
the number of worker processes.  Falls back to the plain-C oracle chain when the reference build is absent."""
import multiprocessing as mp
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_SEEDS = [101, 103, 107, 109, 113, 127, 131, 137, 139, 149, 151, 157, 163, 167, 173, 179]  # CSimulate.cpp:11
N, K = 17664, 14592


def _run_one(args):
    method, lut, mod, il, eb, scale, max_iter, blocks, seed = args
    for p in ("mod-interleaveavx_multithreads-faid_b200", "oracle", "tests"):
        sys.path.insert(0, str(ROOT / p))
    import llrgen
    import pyoracle
    orc = pyoracle.Oracle()
    cfg = orc.default_config(method, lut)
    cfg.mod_type, cfg.interleave_mod_type, cfg.scale, cfg.max_iteration = mod, il, scale, max_iter
    cw = llrgen.golden_codeword()
    sigma = orc.sigma(eb, mod)
    ef = eb_bits = lt3 = 0
    if pyoracle.ref_available("faid3"):
        ref = pyoracle.Ref("faid3")
        sim = pyoracle.RefSim(ref, cfg, seed=seed)
        sim.set_codeword(cw)
        for _ in range(blocks):
            sim.noise_block(sigma, cfg.scale)
            _, st, _ = sim.decode_and_count(method)
            ef += int(st[0]); eb_bits += int(st[1]); lt3 += int(st[2])
        kind = "reference"
    else:
        tx = np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)
        modseq = orc.modulate(tx, mod, il)
        state = np.array([seed, seed, seed], dtype=np.uint64)
        info = np.tile(cw[:K], 32).astype(np.int8)
        for _ in range(blocks):
            sym, state = orc.awgn(modseq, np.float32(sigma / np.sqrt(2)), state)
            _, deint = orc.demodulate(sym, mod, il)
            dec, _ = orc.decode(cfg, orc.quantize(deint, cfg.scale)[None, :])
            st = orc.calc_errors(info, dec[0])
            ef += int(st[0]); eb_bits += int(st[1]); lt3 += int(st[2])
        kind = "port"
    return ef, eb_bits, lt3, 32 * blocks, kind


_pool = None


def pool():
    global _pool
    if _pool is None:
        _pool = mp.get_context("spawn").Pool(min(16, os.cpu_count() or 1))
    return _pool


def reference_points(points, n_seeds=8):
    """points: list of dict(method, lut, mod, il, eb, scale, max_iter, blocks_per_seed) -> list of dict(fer, ber, frames, ...)"""
    jobs, index = [], []
    for i, p in enumerate(points):
        for s in REF_SEEDS[:n_seeds]:
            jobs.append((p["method"], p.get("lut", -1), p["mod"], p["il"], p["eb"], p["scale"], p.get("max_iter", 6), p["blocks_per_seed"], s))
            index.append(i)
    res = pool().map(_run_one, jobs, chunksize=1)
    out = [dict(error_frames=0, error_bits=0, lt3=0, frames=0) for _ in points]
    for i, (ef, ebits, lt3, n, kind) in zip(index, res):
        o = out[i]
        o["error_frames"] += ef; o["error_bits"] += ebits; o["lt3"] += lt3; o["frames"] += n; o["kind"] = kind
    for o in out:
        o["fer"] = o["error_frames"] / o["frames"]
        o["ber"] = o["error_bits"] / (o["frames"] * K)
    return out


if __name__ == "__main__":
    # operating-point scan used to choose the gate's Eb/N0 values: python tests/ref_chain.py
    import json
    pts = []
    for method, mod, il, scale, ebs in ((0, 2, 1, 13.0, (3.6, 3.85)), (2, 2, 1, 13.0, (3.6, 3.8)), (5, 2, 1, 12.5, (3.55, 3.7)),
                                        (4, 4, 4, 13.0, (7.25, 7.45)), (4, 6, 1, 13.0, (12.4, 12.85)), (4, 6, 6, 13.0, (12.3, 12.7))):
        for eb in ebs:
            pts.append(dict(method=method, mod=mod, il=il, eb=eb, scale=scale, blocks_per_seed=int(sys.argv[1]) if len(sys.argv) > 1 else 10))
    for p, r in zip(pts, reference_points(pts)):
        print(json.dumps({**p, **r}))
