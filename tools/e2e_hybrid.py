#!/usr/bin/env python3
"""Hybrid host path (LDPC_B200_HYBRID = number of direct slots) against all-staged, as a function of staged slots and chunk size.
    python tools/e2e_hybrid.py [groups=2048]"""
import os
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import ldpc_b200, llrgen
N, K = 17664, 14592
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_in.array[:] = np.tile(base, (G // 8, 1))
ref = None
for hyb in (0, 1, 2, 3, 4):
    for chunk, streams in ((64, 4), (64, 8), (128, 4), (128, 8), (32, 8)):
        os.environ.pop("LDPC_B200_HYBRID", None)
        if hyb:
            os.environ["LDPC_B200_HYBRID"] = str(hyb)
        cfg = ldpc_b200.default_config(0, -1)
        cfg.chunk_groups, cfg.n_streams = chunk, streams
        with ldpc_b200.Decoder(cfg) as dec:
            for _ in range(2):
                dec.decode(h_in.array, h_out.array)
            t0 = time.perf_counter(); R = 4
            for _ in range(R):
                dec.decode(h_in.array, h_out.array)
            dt = (time.perf_counter() - t0) / R
            r = dec.last_routing()
        if ref is None:
            ref = h_out.array[:8].copy()
        ok = bool((h_out.array[:8] == ref).all())
        print(f"direct slots {hyb} chunk {chunk:4d} staged slots {streams}: {G*32*K/dt/1e9:6.2f} Gbit/s  staged {r['staged_chunks']} direct {r['direct_chunks']} ok {ok}", flush=True)
