#!/usr/bin/env python3
"""A/B timing of decode kernels (CUDA events inside the library), LLRs resident.  One process per library variant:
    LDPC_B200_LIB=build/variants/x.so python tools/nms_ab.py [methods=0] [groups=1024] [ebn0=3.6] [skews=0,...]
Prints one line per (method, skew): kernel ms (mean / min of R launches) and decoded info Gbit/s."""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np, torch
import ldpc_b200, llrgen
N, K = 17664, 14592
methods = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
eb = float(sys.argv[3]) if len(sys.argv) > 3 else 3.6
skews = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
tag = os.environ.get("LDPC_B200_LIB", "default")
base, cw = llrgen.qpsk_llr_groups(8, eb, seed=3)
fix = torch.from_numpy(np.tile(base, (G // 8, 1))).cuda()
out = torch.empty_like(fix)
ref = None
for m in methods:
    cfg = ldpc_b200.default_config(m, -1); cfg.chunk_groups = G; cfg.n_streams = 1
    with ldpc_b200.Decoder(cfg) as dec:
        for sk in skews:
            os.environ["LDPC_B200_SKEW_NS"] = str(sk)
            for _ in range(3): dec.decode(fix, out)
            ts, fs = [], []
            for _ in range(10):
                dec.decode(fix, out)
                a, b = dec.last_timing_detail(); ts.append(a); fs.append(b)
            h = int(out[:8].to(torch.int64).sum().item())
            ts = np.array(ts); fs = np.array(fs)
            tot = ts.mean() + fs.mean()
            print(f"{Path(tag).name} method {m} skew {sk}: decode {ts.mean():.3f} ms (min {ts.min():.3f}) finalize {fs.mean():.3f} ms -> {G*32*K/(tot*1e-3)/1e9:.2f} Gbit/s  chk {h}", flush=True)
