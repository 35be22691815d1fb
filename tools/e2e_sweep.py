"""e2e throughput of ldpc_b200_decode with pinned host buffers as a function of chunk size / stream count."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np, torch
import ldpc_b200, llrgen
N, K = 17664, 14592
G = 1024
base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_in.array[:] = np.tile(base, (G // 8, 1))
for chunk in (32, 64, 128, 256):
    for ns in (2, 3, 4, 6):
        cfg = ldpc_b200.default_config(0, -1); cfg.chunk_groups = chunk; cfg.n_streams = ns
        with ldpc_b200.Decoder(cfg) as dec:
            for _ in range(2): dec.decode(h_in.array, h_out.array)
            torch.cuda.synchronize(); t0 = time.perf_counter(); R = 8
            for _ in range(R): dec.decode(h_in.array, h_out.array)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / R
        print(f"chunk {chunk:4d} streams {ns}: {dt*1e3:6.2f} ms/step  {G*32*K/dt/1e9:6.2f} Gbit/s", flush=True)
