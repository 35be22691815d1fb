#!/usr/bin/env python3
"""Build libldpc_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a.

    python build.py [--force] [-v] [--out=PATH] [-DNAME=VALUE ...] [--kinds=0,1,...]

The seven decode_pair_kernel kinds are ~100 KB of fully unrolled SASS each, so every kind is its own translation unit
(csrc/decode_inst.cu compiled with -DLDPC_INST_KIND=k) and the objects are built in parallel; the C-ABI, the finalize /
frame kernels (csrc/ldpc_b200.cu) and the host staging (csrc/host_pack.cpp) are two more objects.  An object is rebuilt
when any source under csrc/ or include/ is newer than it or when the compiler flags changed (stamp file)."""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
OUT = HERE / "lib" / "libldpc_b200.so"
CSRC = HERE / "csrc"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
KINDS = list(range(7))  # KIND_NMS .. KIND_FAID_ER (decode_kernels.cuh)


def deps():
    return [p for p in CSRC.glob("*") if p.is_file()] + list((ROOT / "include").glob("*.h"))


def needs_build(out=None):
    out = Path(out) if out else OUT
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(p.stat().st_mtime > t for p in deps())


def _flags(extra, verbose):
    return ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
            "-Xptxas", "-v" if verbose else "-warn-spills", "-I", str(ROOT / "include"), "-I", str(CSRC)] + list(extra)


def _compile(job):
    name, src, flags, obj, stamp, key, force, verbose = job
    newest = max(p.stat().st_mtime for p in deps())
    if not force and obj.exists() and obj.stat().st_mtime >= newest and stamp.exists() and stamp.read_text() == key:
        return name, 0, ""
    cmd = [NVCC, "-c", "-o", str(obj)] + flags + [str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        stamp.write_text(key)
    return name, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, out=None, extra=(), kinds=None, jobs=None):
    """extra: additional nvcc flags (e.g. -DLDPC_CV_SMEM_LAYERS=5 for A/B experiments written to another `out`).
    kinds: subset of kernel kinds to (re)compile -- the others' launchers are still linked from their existing objects."""
    out = Path(out) if out else OUT
    out.parent.mkdir(parents=True, exist_ok=True)
    tag = hashlib.sha1((" ".join(extra) + str(out)).encode()).hexdigest()[:10] if (extra or out != OUT) else "default"
    # default build: objects in-tree next to the library (they travel to the GPU box, where build.py then finds everything
    # up to date); experiment variants keep theirs beside their own output (build/variants/obj is gpurun-ignored)
    objdir = (HERE / "lib" / "obj" / tag) if tag == "default" else (out.parent / "obj" / tag)
    objdir.mkdir(parents=True, exist_ok=True)
    flags = _flags(extra, verbose)
    key = " ".join(flags)
    todo = []
    for k in KINDS:
        f = flags + [f"-DLDPC_INST_KIND={k}"]
        todo.append((f"kind{k}", CSRC / "decode_inst.cu", f, objdir / f"decode_kind{k}.o", objdir / f"decode_kind{k}.stamp", key,
                     force and (kinds is None or k in kinds), verbose))
    todo.append(("api", CSRC / "ldpc_b200.cu", flags, objdir / "ldpc_b200.o", objdir / "ldpc_b200.stamp", key, force, verbose))
    todo.append(("host_pack", CSRC / "host_pack.cpp", flags, objdir / "host_pack.o", objdir / "host_pack.stamp", key, force, verbose))
    with ThreadPoolExecutor(max_workers=jobs or min(len(todo), os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile, todo))
    bad = [r for r in results if r[1] != 0]
    for name, rc, log in results:
        if log and (verbose or rc != 0 or "spill" in log.lower()):
            sys.stderr.write(f"--- {name} ---\n{log}")
    if bad:
        raise RuntimeError("nvcc failed: " + ", ".join(r[0] for r in bad))
    objs = [str(j[3]) for j in todo]
    if not out.exists() or any(Path(o).stat().st_mtime > out.stat().st_mtime for o in objs) or force:
        r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out)] + objs + ["-ldl", "-lpthread"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    # static instruction mix of the kernels just linked (bench.py's pipe roofline reads it and checks the hash against the
    # library it loaded, so a stale mix cannot be reported)
    if out == OUT:
        write_sass_mix(out)
    return out


def write_sass_mix(lib=None, force=False):
    lib = Path(lib) if lib else OUT
    dst = lib.parent / "sass_mix.json"
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_mix
    if not force and dst.exists() and dst.stat().st_mtime >= lib.stat().st_mtime:
        try:
            import json
            if json.loads(dst.read_text()).get("lib_sha256") == sass_mix.lib_sha256(lib):
                return dst
        except Exception:
            pass
    import json
    dst.write_text(json.dumps(sass_mix.mix_of(lib), indent=1) + "\n")
    return dst


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a.startswith("-D") or a.startswith("-maxrregcount")]
    for a in sys.argv[1:]:  # --ptxas=-regUsageLevel=7 -> -Xptxas -regUsageLevel=7
        if a.startswith("--ptxas="):
            extra += ["-Xptxas", a.split("=", 1)[1]]
    out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    kinds = next(([int(x) for x in a.split("=", 1)[1].split(",")] for a in sys.argv[1:] if a.startswith("--kinds=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=out, extra=extra, kinds=kinds))
