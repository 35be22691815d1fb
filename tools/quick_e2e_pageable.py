"""e2e of ldpc_b200_decode with PAGEABLE caller arrays (plain numpy = malloc, what the reference's fixInput / decodedBits
are) vs pinned ones, with the host staging on (default) and off (LDPC_B200_HOST_THREADS=0)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import ldpc_b200, llrgen
N, K, G = 17664, 14592, 1024
base = llrgen.qpsk_llr_groups(8, 3.6, seed=3)[0]
pin_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8); pin_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
pin_in.array[:] = np.tile(base, (G // 8, 1))
pg_in = np.ascontiguousarray(pin_in.array.copy()); pg_out = np.empty_like(pg_in)
for threads in ("default", "0"):
    if threads == "0":
        os.environ["LDPC_B200_HOST_THREADS"] = "0"
    cfg = ldpc_b200.default_config(0, -1)
    with ldpc_b200.Decoder(cfg) as dec:
        for name, a, b in (("pinned", pin_in.array, pin_out.array), ("pageable", pg_in, pg_out)):
            for _ in range(2): dec.decode(a, b)
            t0 = time.perf_counter(); R = 3
            for _ in range(R): dec.decode(a, b)
            dt = (time.perf_counter() - t0) / R
            print(f"host threads {threads:7s} {name:8s}: {G*32*K/dt/1e9:.2f} Gbit/s  staging {dec.host_staging()['threads']}", flush=True)
        assert (pg_out == pin_out.array).all()
