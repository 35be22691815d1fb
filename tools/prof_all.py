"""Launch every kernel of the engine once on a mid-size workload, for `ncu --set full` (one capture per kernel kind).
Usage: ncu --set full --clock-control none -k regex:'decode_pair|finalize|generate|encode|count_errors|info_bits|quantize|group_hist' \
           -f -o gpurun_out/prof_all python tools/prof_all.py [groups]"""
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import torch
import ldpc_b200
import llrgen

N, K = 17664, 14592
G = int(sys.argv[1]) if len(sys.argv) > 1 else 296
cw = llrgen.golden_codeword()
tx_one = np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)
d_tx = torch.from_numpy(tx_one).cuda().repeat(G, 1)
d_info = torch.from_numpy(np.tile(cw[:K], 32).astype(np.int8)).cuda().repeat(G, 1)


def run(method, lut=-1, general_faid=False, eb=3.6, mod=2, il=1, tag=""):
    if general_faid:
        os.environ["LDPC_B200_NO_FAID_FAST"] = "1"
    else:
        os.environ.pop("LDPC_B200_NO_FAID_FAST", None)
    cfg = ldpc_b200.default_config(method, lut)
    cfg.mod_type, cfg.interleave_mod_type = mod, il
    cfg.chunk_groups = G
    cfg.n_streams = 1
    with ldpc_b200.Decoder(cfg) as dec:
        fix = dec.generate(d_tx, eb, 101, 0, G)                 # generate_kernel
        out = dec.decode(fix)                                   # decode_pair_kernel<kind> + finalize_kernel(bf mode)
        c = dec.count_errors(d_info, out)                       # count_errors_kernel
        print(tag or f"method {method}", "FER", c[1] / c[0], dec.last_timing_detail())
        return dec, fix


run(0)
run(1)
run(2)
run(2, general_faid=True, tag="method 2 (general LUT path)")
run(3)
run(4)
run(5)
run(5, general_faid=True, tag="method 5 (general LUT path)")
run(4, eb=12.0, mod=6, il=6, tag="method 4, 64-QAM")
os.environ.pop("LDPC_B200_NO_FAID_FAST", None)
# the remaining frame kernels: info bits + encoder (simulate with random info), fused producer, quantiser, demapper
cfg = ldpc_b200.default_config(0, -1)
cfg.chunk_groups = G
with ldpc_b200.Decoder(cfg) as dec:
    print("simulate (random info, fused producer)", dec.simulate(3.6, 5, 0, G)[:4])
    x = torch.randn(G * 32 * N // 8, device="cuda")
    dec.quantize(x)
    sym = torch.randn(8, 2 * 32 * N // 2, device="cuda")
    dec.demap(sym)
torch.cuda.synchronize()
