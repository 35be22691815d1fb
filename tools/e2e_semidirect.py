#!/usr/bin/env python3
"""Host-buffer call: direct slots whose OUTPUT still travels as bits (LDPC_B200_HYBRID_OUT_BITS=1) against the shipped hybrid
(direct slots copy both arrays as they are) and against input-as-it-is / output-as-bits for every chunk.
    python tools/e2e_semidirect.py [groups=2048]"""
import os
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import ldpc_b200, llrgen
N, K = 17664, 14592
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
h_in.array[:] = np.tile(base, (G // 8, 1))
ref = None


def run(tag, env, chunk, streams):
    global ref
    for k in ("LDPC_B200_HYBRID", "LDPC_B200_HYBRID_OUT_BITS", "LDPC_B200_HYBRID_CHUNK", "LDPC_B200_STAGE_IN", "LDPC_B200_STAGE_OUT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cfg = ldpc_b200.default_config(0, -1)
    if chunk:
        cfg.chunk_groups, cfg.n_streams = chunk, streams
    with ldpc_b200.Decoder(cfg) as dec:
        for _ in range(2):
            dec.decode(h_in.array, h_out.array)
        h_out.array[:] = 0x55
        best = 1e9
        tot = 0.0
        R = 5
        for _ in range(R):
            t0 = time.perf_counter()
            dec.decode(h_in.array, h_out.array)
            dt = time.perf_counter() - t0
            best = min(best, dt); tot += dt
        r = dec.last_routing()
    if ref is None:
        ref = h_out.array.copy()
    ok = bool((h_out.array == ref).all())
    print(f"{tag:34s} chunk {chunk or 0:4d} x {streams or 0}: mean {G*32*K/(tot/R)/1e9:6.2f} best {G*32*K/best/1e9:6.2f} Gbit/s  staged {r['staged_chunks']} direct {r['direct_chunks']} ok {ok}", flush=True)


run("shipped default", {}, None, None)
for rep in range(2):
    run("shipped default", {}, None, None)
    for nd in (1, 2):
        for chunk, streams in ((64, 6), (64, 8), (96, 6), (48, 8), (128, 6)):
            for dchunk in sorted({chunk, chunk // 2, 32}, reverse=True):
                run(f"hybrid {nd} direct x {dchunk}, out bits", {"LDPC_B200_HYBRID": str(nd), "LDPC_B200_HYBRID_OUT_BITS": "1", "LDPC_B200_HYBRID_CHUNK": str(dchunk)}, chunk, streams)
    run("hybrid 1 direct (raw both)", {"LDPC_B200_HYBRID": "1"}, 128, 4)
    run("all staged", {"LDPC_B200_HYBRID": "0"}, 128, 6)
