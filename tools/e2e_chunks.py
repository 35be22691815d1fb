#!/usr/bin/env python3
"""Host-buffer call with full staging as a function of chunk size / slot count (and, per process, LDPC_B200_PACK_NT).
    [LDPC_B200_PACK_NT=0] python tools/e2e_chunks.py [groups=2048]"""
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
for chunk, streams in ((128, 3), (128, 5), (64, 4), (32, 6), (16, 8)):
    cfg = ldpc_b200.default_config(0, -1)
    cfg.chunk_groups, cfg.n_streams = chunk, streams
    with ldpc_b200.Decoder(cfg) as dec:
        for _ in range(2):
            dec.decode(h_in.array, h_out.array)
        t0 = time.perf_counter(); R = 5
        for _ in range(R):
            dec.decode(h_in.array, h_out.array)
        dt = (time.perf_counter() - t0) / R
    print(f"PACK_NT={os.environ.get('LDPC_B200_PACK_NT', '1')} chunk {chunk:4d} slots {streams}: {G*32*K/dt/1e9:6.2f} Gbit/s", flush=True)
