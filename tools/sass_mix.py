#!/usr/bin/env python3
"""Static instruction mix of the message-passing kernels, from the SASS of the built library.

    python tools/sass_mix.py [--lib path/to/libldpc_b200.so] [--out profiles/sass_mix_rNN.json]

For every decode_pair_kernel<KIND, MONO> instantiation: finds the iteration loop (the longest backward branch),
counts its instructions per opcode and assigns them to the sm_100a execution pipes as measured by
profiles/microbench (VIADD / VIADD.16 / IMAD issue beside the ALU-pipe ops, i.e. on the FMA side; everything
bit-wise, min/max, PRMT, SHF, VABSDIFF4 on the ALU pipe at 64 lanes/clk/SM) and cross-checked against ncu's
sm__inst_executed_pipe_* for the NMS kernel (profiles/r01_nms_v2_ncu_full.md: 14.0 ALU-pipe instructions per
pair-edge predicted and measured).  bench.py uses the per-edge ALU-pipe count to report the pipe roofline.
"""
import argparse
import json
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
EDGES = 275  # circulants = edges per check-row-thread per iteration

ALU = ("LOP3", "VIMNMX3", "HMNMX2", "VIADDMNMX", "VIMNMX", "VIMNMX3", "VABSDIFF4", "VABSDIFF", "SHF", "PRMT", "ISETP", "VOTE", "LEA", "PLOP3", "SEL", "POPC", "FLO", "BREV", "IABS")
FMA = ("IMAD", "VIADD", "IADD3", "IADD", "MOV", "FFMA", "FADD", "FMUL", "HADD2", "HFMA2", "IDP")
LSU = ("LDS", "STS", "LDL", "STL", "LDG", "STG", "LD", "ST", "ATOMS", "ATOMG", "ATOM", "RED", "LDSM")


def pipe_of(op):
    base = op.split(".")[0]
    if base in ALU:
        return "alu"
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base.startswith("U") or base in ("S2UR", "R2UR", "LDCU"):
        return "uniform"
    return "other"


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
        if m and cur:
            addr = int(m.group(1), 16)
            ins = m.group(2)
            ins = re.sub(r"^@!?U?P\w+\s+", "", ins)
            out[cur].append((addr, ins))
    return out


def analyse(instrs):
    """The ITERATION loop: the union of the backward branches whose body holds at least one BAR per layer (12) but not every BAR
    of the kernel (ptxas gives the loop several back edges -- early-stop paths, rotated headers -- and the enclosing work loop of
    the persistent variant, which also holds the load phase's barriers, must stay out).  Conditional blocks inside the range
    (hard-decision snapshots of converged frames) are counted, so the figure is an upper bound for the common path."""
    loops = []
    for addr, ins in instrs:
        m = re.match(r"BRA(?:\.\w+)*\s+.*?(0x[0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                loops.append((tgt, addr))
    all_bars = [a for a, i in instrs if i.startswith("BAR")]
    cands = []
    for lo, hi in loops:
        bars = sum(1 for a in all_bars if lo <= a <= hi)
        if bars >= 12:
            cands.append((lo, hi, bars))
    if not cands:
        return None
    inner = [c for c in cands if c[2] < len(all_bars)]
    if inner:
        best = (min(c[0] for c in inner), max(c[1] for c in inner))
    else:
        best = min(((c[0], c[1]) for c in cands), key=lambda r: r[1] - r[0])
    body = [ins for addr, ins in instrs if best[0] <= addr <= best[1]]
    ops = Counter(ins.split()[0] for ins in body)
    pipes = Counter()
    for op, n in ops.items():
        pipes[pipe_of(op)] += n
    return {"loop_instructions": len(body), "per_edge": {k: round(v / EDGES, 3) for k, v in sorted(pipes.items())},
            "per_edge_total": round(len(body) / EDGES, 3),
            "top_opcodes": {k: v for k, v in ops.most_common(16)}}


def lib_sha256(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


NAMES = {"Li0ELb1E": "NMS", "Li0ELb0E": "NMS_general_scale", "Li1ELb1E": "OMS", "Li1ELb0E": "OMS_nomono", "Li2ELb1E": "FAID", "Li3ELb1E": "FAID_EF",
         "Li4ELb1E": "FAID_M", "Li5ELb1E": "FAID_EF_M", "Li6ELb1E": "FAID_ER"}


def mix_of(lib):
    """-> {"lib_sha256": ..., "kinds": {name: analysis}} for the decode_pair_kernel instantiations of `lib`."""
    res = {}
    for fn, instrs in kernels(lib).items():
        if "decode_pair_kernel" not in fn:
            continue
        key = next((v for k, v in NAMES.items() if k in fn), fn)
        r = analyse(instrs)
        if r:
            res[key] = r
    return {"lib_sha256": lib_sha256(lib), "lib": str(lib), "kinds": res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=str(ROOT / "mod-interleaveavx_multithreads-faid_b200" / "lib" / "libldpc_b200.so"))
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    txt = json.dumps(mix_of(a.lib), indent=1)
    if a.out:
        Path(a.out).write_text(txt + "\n")
    print(txt)


if __name__ == "__main__":
    sys.exit(main())
