#!/usr/bin/env python3
"""Stand-alone speed of the host staging passes (csrc/host_pack.cpp) on this box, no GPU activity: nibble packing, bit expansion,
the fused pass, and a streaming copy as the yardstick of what the host's memory system gives the same threads.
    python tools/hostpack_bench.py [groups=512] [threads=all]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
csrc = ROOT / "mod-interleaveavx_multithreads-faid_b200" / "csrc"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else len(os.sched_getaffinity(0))
so = Path(tempfile.mkdtemp()) / "libhp.so"
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", f"-I{csrc}", f"-I{ROOT / 'include'}",
                str(ROOT / "tools" / "probe" / "host_pack_capi.cpp"), str(csrc / "host_pack.cpp"), "-o", str(so)], check=True)
lib = C.CDLL(str(so))
lib.hp_bench.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
N, K = 17664, 14592
for thr in sorted({T, max(1, T // 2), max(1, T // 4)}, reverse=True):
    out = (C.c_double * 4)()
    assert lib.hp_bench(G, thr, 5, out) == 0
    f = G * 32
    print(f"{thr:3d} threads, {G} groups: pack {out[0]*1e3:7.2f} ms = {f*K/out[0]/1e9:6.1f} Gbit/s info ({f*N*1.5/out[0]/1e9:6.1f} GB/s of DRAM traffic) | "
          f"expand {out[1]*1e3:7.2f} ms = {f*K/out[1]/1e9:6.1f} Gbit/s ({f*N*1.125/out[1]/1e9:6.1f} GB/s) | "
          f"fused {out[2]*1e3:7.2f} ms = {f*K/out[2]/1e9:6.1f} Gbit/s ({f*N*2.625/out[2]/1e9:6.1f} GB/s) | "
          f"streaming copy {f*N*2/out[3]/1e9:6.1f} GB/s", flush=True)
