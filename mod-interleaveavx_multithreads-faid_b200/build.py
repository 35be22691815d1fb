#!/usr/bin/env python3
"""Build libldpc_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a.  Usage: python build.py [--force]"""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
OUT = HERE / "lib" / "libldpc_b200.so"
SRCS = [HERE / "csrc" / "ldpc_b200.cu", HERE / "csrc" / "host_pack.cpp"]  # the .cpp goes straight to g++ (AVX-512 bodies behind a run-time check)
DEPS = list((HERE / "csrc").glob("*")) + list((ROOT / "include").glob("*.h"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build():
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(p.stat().st_mtime > t for p in DEPS)


def build(force=False, verbose=False, out=None, extra=()):
    """extra: additional nvcc flags (e.g. -DLDPC_ADDR_HI=0 for A/B experiments written to another `out`)."""
    global OUT
    if out is not None:
        OUT = Path(out)
        force = True
    if not force and not needs_build():
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-I", str(ROOT / "include"), "-I", str(HERE / "csrc"), "-o", str(OUT)] + list(extra) + [str(s) for s in SRCS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D") or a.startswith("-maxrregcount")]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out, extra=extra))
