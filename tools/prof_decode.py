"""Small fixed workload for ncu: one decode of G groups, methods given on the command line."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np, torch
import ldpc_b200, llrgen
G = int(sys.argv[2]) if len(sys.argv) > 2 else 74
methods = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
base, cw = llrgen.qpsk_llr_groups(2, 3.6, seed=3)
fix = torch.from_numpy(np.tile(base, (G // 2, 1))).cuda()
out = torch.empty_like(fix)
for m in methods:
    cfg = ldpc_b200.default_config(m, -1); cfg.chunk_groups = G; cfg.n_streams = 1
    with ldpc_b200.Decoder(cfg) as dec:
        for _ in range(3):
            dec.decode(fix, out)
        print(m, dec.last_timing(), dec.last_timing_detail())
