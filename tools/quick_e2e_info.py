"""e2e of ldpc_b200_decode (host int8 arrays) with and without the per-group outputs (BF iterations, executed iterations,
convergence iteration) -- the form a CSimulate-style caller uses (it histograms the returned BFiter)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import ldpc_b200, llrgen
N, K, G = 17664, 14592, 1024
base = llrgen.qpsk_llr_groups(8, 3.6, seed=3)[0]
h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8); h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_in.array[:] = np.tile(base, (G // 8, 1))
for method in (0, 2, 4):
    cfg = ldpc_b200.default_config(method, -1); cfg.chunk_groups, cfg.n_streams = 128, 3
    with ldpc_b200.Decoder(cfg) as dec:
        for want in (False, True):
            for _ in range(2): dec.decode(h_in.array, h_out.array, want_info=want)
            t0 = time.perf_counter(); R = 4
            for _ in range(R): dec.decode(h_in.array, h_out.array, want_info=want)
            dt = (time.perf_counter() - t0) / R
            print(f"method {method} want_info={want}: {G*32*K/dt/1e9:.2f} Gbit/s", flush=True)
