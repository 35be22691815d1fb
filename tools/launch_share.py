#!/usr/bin/env python3
"""Kernel shares of a `bench.py` step from the ncu launch list (gpu__time_duration.sum, cold-cache, serialised) next to
the CUDA-event shares bench.py reports.  Usage: tools/launch_share.py launches.csv bench.json out.md"""
import csv
import json
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if r and r[0] == "ID":
        hdr = r
        break
idx = {h: i for i, h in enumerate(hdr)}
dur = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("ldpc::", "")
    if "at::native" in r[idx["Kernel Name"]] or "elementwise" in name or "nccl" in name.lower():
        name = "(torch / other)"
    dur[name] += float(r[idx["Metric Value"]].replace(",", "")) * (1e-3 if r[idx["Metric Unit"]] in ("ns", "nsecond") else 1.0)
    cnt[name] += 1
bench = json.load(open(sys.argv[2]))
km = bench["kernel_ms_per_step"]
tot_evt = km["decode_pair_kernel"] + km["finalize_kernel"]
own = {k: v for k, v in dur.items() if "decode_pair_kernel<0" in k or k.startswith("finalize_kernel")}
# the bench's device-resident steps are the launches of decode_pair_kernel<0,..> with the full grid + the finalize after each
out = ["ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu --no-methods` (one row per kernel name; durations are",
       "cold-cache and serialised, so only the SHARES are comparable with the CUDA-event timing of the un-profiled run).", "",
       "| kernel | launches | total us | share of (decode + finalize) |", "|---|---|---|---|"]
tot_own = sum(own.values())
for k, v in sorted(dur.items(), key=lambda kv: -kv[1]):
    share = f"{100 * v / tot_own:.1f} %" if k in own else ""
    out.append(f"| `{k}` | {cnt[k]} | {v:.1f} | {share} |")
out += ["", "For DecodeMethod 0 the device-resident step is one `decode_pair_kernel` launch (it writes decodedBits itself); the "
        "`finalize_kernel`, `generate_i1_kernel`, `count_errors_kernel` and `group_hist_kernel` launches belong to the "
        "`ldpc_b200_simulate` rounds and to the scoring of the bench, outside the `value` timing."]
out += ["", f"CUDA events in the un-profiled bench (`kernel_ms_per_step`): decode_pair_kernel {km['decode_pair_kernel']:.3f} ms = "
        f"{100 * km['decode_pair_kernel'] / tot_evt:.1f} %, finalize_kernel {km['finalize_kernel']:.3f} ms = {100 * km['finalize_kernel'] / tot_evt:.1f} %."]
open(sys.argv[3], "w").write("\n".join(out) + "\n")
print("\n".join(out))
