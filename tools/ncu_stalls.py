#!/usr/bin/env python3
"""Warp-stall sampling of one kernel from an `ncu --set full --import-source on` report: totals per stall reason, per SASS
opcode, and the instructions that collect the most samples of a chosen reason.
Usage: tools/ncu_stalls.py gpurun_out/x.ncu-rep profiles/<name>.md [reason=long_sb]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    reason = "stall_" + (sys.argv[3] if len(sys.argv) > 3 else "long_sb")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], [r for r in rows[2:] if len(r) >= len(rows[1])]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    tot, byop, cnt, execs = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        cnt[op] += f(r, "# Samples")
        execs[op] += f(r, "Instructions Executed")
        for s in stalls:
            tot[s] += f(r, s)
            byop[op][s] += f(r, s)
    T, E = sum(cnt.values()), sum(execs.values())
    lines = [f"Warp-stall sampling, `{kernel}` (`ncu --set full --import-source on --clock-control none`, {int(T)} samples, {int(E)} warp-instructions).", "",
             "| stall reason | share of samples |", "|---|---|"]
    lines += [f"| {k[6:]} | {v / T:.3f} |" for k, v in tot.most_common(12)]
    lines += ["", "| opcode | share of samples | share of executed instructions | top reasons (samples per sample of this opcode) |", "|---|---|---|---|"]
    for op, n in cnt.most_common(16):
        top = ", ".join(f"{k[6:]} {v / n:.2f}" for k, v in byop[op].most_common(4))
        lines.append(f"| {op} | {n / T:.3f} | {execs[op] / E:.3f} | {top} |")
    worst = sorted(((f(r, reason), r[ix["Source"]].strip()) for r in data), reverse=True)[:12]
    lines += ["", f"Instructions with the most `{reason[6:]}` samples (total {int(tot[reason])}, {tot[reason] / T:.3f} of all):", ""]
    lines += [f"* {int(v)}  `{s}`" for v, s in worst if v > 0]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
