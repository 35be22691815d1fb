"""Kernel time of DecodeMethod 2 with EF_ELIMINATION 0 / 1 / 2 (fast LUT kinds vs the erasure kind) on one B200."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np, torch
import ldpc_b200, llrgen
N, K = 17664, 14592
G = 1024
base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
fix = torch.from_numpy(np.tile(base, (G // 8, 1))).cuda()
out = torch.empty_like(fix)
for ef in (0, 1, 2):
    cfg = ldpc_b200.default_config(2, -1); cfg.chunk_groups = G
    cfg.ef_elimination = ef
    if ef:
        cfg.ef_floor_err_count, cfg.ef_floor_iter_thresh = (50 if ef == 1 else 20), 6
    with ldpc_b200.Decoder(cfg) as dec:
        for _ in range(2): dec.decode(fix, out)
        kd = kf = 0.0
        R = 5
        for _ in range(R):
            dec.decode(fix, out); a, b = dec.last_timing_detail(); kd += a; kf += b
        fr = G * 32
        print(f"EF_ELIMINATION {ef}: decode {kd/R:.2f} ms + finalize {kf/R:.2f} ms -> {fr*K/((kd+kf)/R*1e-3)/1e9:.2f} info Gbit/s (kernel-only)")
