#!/usr/bin/env python3
"""Host-buffer call ldpc_b200_decode(): throughput as a function of staging mode, thread count, chunk size and NUMA placement.
    python tools/e2e_exp.py [groups=1024] [quick]
One line per configuration; the decoded bits of every configuration are checked against the first one."""
import itertools
import os
import subprocess
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np
import torch
import ldpc_b200, llrgen
N, K = 17664, 14592
G = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
quick = len(sys.argv) > 2
for cmd in (["lscpu"], ["numactl", "-H"], ["nvidia-smi", "topo", "-m"]):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
        keep = [l for l in out.splitlines() if any(k in l for k in ("Model name", "Socket", "Core(s)", "Thread(s)", "NUMA", "CPU(s):", "node", "GPU", "L3"))]
        print("\n".join(keep[:24]), flush=True)
    except Exception as e:
        print(cmd, "unavailable:", e)
print("affinity cpus:", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), flush=True)
base, cw = llrgen.qpsk_llr_groups(8, 3.6, seed=3)
ref_out = None


def run(env, chunk, streams, numa, label):
    global ref_out
    keys = ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN", "LDPC_B200_NUMA", "LDPC_B200_HYBRID")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["LDPC_B200_NUMA"] = "1" if numa else "0"
    h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8)     # placed by ldpc_b200_host_alloc under the current NUMA setting
    h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
    h_in.array[:] = np.tile(base, (G // 8, 1))
    cfg = ldpc_b200.default_config(0, -1)
    cfg.chunk_groups, cfg.n_streams = chunk, streams
    with ldpc_b200.Decoder(cfg) as dec:
        st = dec.host_staging(); pl = dec.host_placement()
        for _ in range(2):
            dec.decode(h_in.array, h_out.array)
        R = 3 if quick else 6
        t0 = time.perf_counter()
        for _ in range(R):
            dec.decode(h_in.array, h_out.array)
        dt = (time.perf_counter() - t0) / R
        st2 = dec.host_staging()
    ok = True
    if ref_out is None:
        ref_out = h_out.array[:16].copy()
    else:
        ok = bool((h_out.array[:16] == ref_out).all())
    print(f"{label:34s} chunk {chunk:4d} streams {streams} numa {int(numa)} (node {pl['numa_node']}, {pl['numa_cpus']} cpus) threads {st['threads']:3d} "
          f"in {int(st['stage_in'])} out {int(st['stage_out'])}: {G*32*K/dt/1e9:6.2f} Gbit/s  {G*32/dt/1e6:5.2f} Mframes/s  "
          f"h2d {st2['last_h2d_bytes']/G/32:.0f} B/frame d2h {st2['last_d2h_bytes']/G/32:.0f} ok {ok}", flush=True)
    h_in.free(); h_out.free()


ncpu = len(os.sched_getaffinity(0))
modes = [("direct", {"LDPC_B200_HOST_THREADS": "0"})]
for thr in sorted({8, 16, 32, ncpu} & set(range(1, ncpu + 1))):
    modes.append((f"staged in+out, {thr} threads", {"LDPC_B200_HOST_THREADS": str(thr), "LDPC_B200_STAGE_IN": "1", "LDPC_B200_STAGE_OUT": "1"}))
modes.append(("default", {}))
modes.append(("bits out only, default threads", {"LDPC_B200_STAGE_IN": "0", "LDPC_B200_STAGE_OUT": "1"}))
for numa in (True, False):
    for label, env in modes:
        for chunk, streams in ((128, 3),) if quick else ((128, 3), (64, 4), (32, 6)):
            run(env, chunk, streams, numa, label)
