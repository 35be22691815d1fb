import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np, torch
import ldpc_b200, llrgen
N, K = 17664, 14592
G = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
methods = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0]
eb = float(sys.argv[3]) if len(sys.argv) > 3 else 3.6
base, cw = llrgen.qpsk_llr_groups(8, eb, seed=3)
fix = torch.from_numpy(np.tile(base, (G // 8, 1))).cuda()
out = torch.empty_like(fix)
for m in methods:
    cfg = ldpc_b200.default_config(m, -1); cfg.chunk_groups = G
    with ldpc_b200.Decoder(cfg) as dec:
        for _ in range(2): dec.decode(fix, out)
        torch.cuda.synchronize(); t0 = time.time()
        R = 5
        kms = 0; kd = 0; kf = 0
        for _ in range(R):
            dec.decode(fix, out); kms += dec.last_timing()[0]; a, b = dec.last_timing_detail(); kd += a; kf += b
        torch.cuda.synchronize(); dt = (time.time() - t0) / R
        fr = G * 32
        print(f"method {m}: {dt*1e3:.2f} ms/step wall, kernel {kms/R:.2f} ms, {fr/dt/1e6:.3f} Mframes/s, {fr*K/dt/1e9:.2f} info Gbps (kernel-only {fr*K/(kms/R*1e-3)/1e9:.2f}) [decode {kd/R:.2f} ms + finalize {kf/R:.2f} ms]")
