#!/usr/bin/env python3
"""ldpc_b200_decode with PAGEABLE caller arrays (what the reference's malloc'd fixInput / decodedBits are) against pinned ones, and
with LDPC_B200_HOST_REGISTER=1 (the library page-locks the caller's arrays on first use).  One process per setting.
    python tools/e2e_pageable.py [groups=2048]"""
import os
import subprocess
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
if len(sys.argv) > 2:  # worker
    for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
        sys.path.insert(0, str(ROOT / p))
    import time
    import numpy as np
    import ldpc_b200, llrgen
    N, K = 17664, 14592
    kind = sys.argv[2]
    base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
    if kind == "pinned":
        a_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8); a_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
        x, y = a_in.array, a_out.array
    else:
        x = np.empty((G, 32 * N), np.int8); y = np.empty((G, 32 * N), np.int8)
    x[:] = np.tile(base, (G // 8, 1)); y[:] = 0
    import re
    def thp():
        return sum(int(v) for v in re.findall(r"AnonHugePages:\s+(\d+) kB", open("/proc/self/smaps").read())) // 1024
    thp0 = thp()
    with ldpc_b200.Decoder(ldpc_b200.default_config(0, -1)) as dec:
        for _ in range(2):
            dec.decode(x, y)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); dec.decode(x, y); ts.append(time.perf_counter() - t0)
        r = dec.last_routing(); st = dec.host_staging()
    print(f"{kind:10s} HOST_REGISTER={os.environ.get('LDPC_B200_HOST_REGISTER', '0')} (huge pages of the process {thp()} MB, decodedBits at 64k+{y.ctypes.data % 64}): mean {G*32*K/(sum(ts)/len(ts))/1e9:6.2f} best {G*32*K/min(ts)/1e9:6.2f} Gbit/s  "
          f"staged {r['staged_chunks']} direct {r['direct_chunks']} threads {st['threads']} h2d {st['last_h2d_bytes']/(G*32):.0f} d2h {st['last_d2h_bytes']/(G*32):.0f} B/frame", flush=True)
    sys.exit(0)
for kind, reg in (("pageable", "0"), ("pageable", "1"), ("pinned", "0"), ("pageable", "0"), ("pageable", "1"), ("pinned", "0")):
    env = dict(os.environ, LDPC_B200_HOST_REGISTER=reg)
    subprocess.run([sys.executable, __file__, str(G), kind], env=env, check=False)
