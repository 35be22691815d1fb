import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "oracle", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import ldpc_b200, pyoracle, llrgen
O = pyoracle.Oracle()
N, K = 17664, 14592
method = int(sys.argv[1]) if len(sys.argv) > 1 else 0
fix, cw = llrgen.qpsk_llr_groups(1, 3.6, seed=5)
for mi in (0, 1, 2, 6):
    cfg = ldpc_b200.default_config(method, -1); cfg.max_iteration = mi
    ocfg = O.default_config(method, -1); ocfg.max_iteration = mi
    with ldpc_b200.Decoder(cfg) as dec:
        out, info = dec.decode(fix, want_info=True)
    ref, infos = O.decode(ocfg, fix)
    d = (out != ref).reshape(32, N)
    print(f"method {method} max_iter {mi}: diff total {int(d.sum())}; per frame {d.sum(1)[:8]}...; per block col (frame0) {d[0].reshape(69,256).sum(1)[:12]} its {info['its_per_group']} {infos[0].iters_executed} bf {info['bf_iters']} {infos[0].bf_iters}")
    if d.sum():
        idx = np.argwhere(d)[:10]
        print("  first diffs (frame, n):", idx.tolist())
