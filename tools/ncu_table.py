#!/usr/bin/env python3
"""One-line-per-launch roofline table from an `ncu --set full` report (read here with `ncu -i`).
Usage: tools/ncu_table.py gpurun_out/prof_all.ncu-rep profiles/r01_all_kernels [hbm_peak_gbs]"""
import csv
import json
import re
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    u = unit.lower()
    scale = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3,
             "ms": 1e3, "nsecond": 1e-3, "second": 1e6}
    return x * scale.get(u, 1.0)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6551.4
    if rep.endswith(".csv"):  # `ncu -i rep --page raw --csv` already run on the GPU box (the .ncu-rep was too big to bring back)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        name = re.sub(r"\(ldpc::\w+\)|\(.*\)$", "", d.get("Kernel Name", "")).replace("ldpc::", "").replace("void ", "").strip()
        k = {"id": int(d.get("ID", 0)), "kernel": name}
        for key, short in COLS:
            if key in d:
                k[short] = to_float(d[key], u[key])
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")}
        k["top_stalls"] = [a for a, _ in sorted(stalls.items(), key=lambda kv: -kv[1])[:3]]
        t = k.get("time_us") or 0
        if t:
            k["dram_gbs"] = ((k.get("dram_rd_MB") or 0) + (k.get("dram_wr_MB") or 0)) * 1e6 / (t * 1e-6) / 1e9
            k["hbm_frac"] = k["dram_gbs"] / peak
        res.append(k)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"ncu --set full --clock-control none, one row per launch (HBM peak used for hbm_frac: {peak} GB/s measured copy bandwidth).\n"
                "Pipe columns are sm__inst_executed_pipe_*.avg.pct_of_peak_sustained_active; the binding one is in **bold**.\n\n")
        f.write("| # | kernel | grid x block | regs | time us | DRAM GB/s | HBM frac | ALU % | FMA % | LSU % | XU % | issue % | occ % | top stalls |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k in res:
            pipes = {p: (k.get(p + "_pct") or 0) for p in ("alu", "fma", "lsu", "xu")}
            top = max(pipes, key=pipes.get)
            hb = (k.get("hbm_frac") or 0) * 100

            def cell(p):
                v = f"{pipes[p]:.1f}"
                return f"**{v}**" if p == top and pipes[p] >= hb else v
            hcell = f"{hb:.1f} %"
            if hb > pipes[top]:
                hcell = f"**{hcell}**"
            f.write(f"| {k['id']} | {k['kernel']} | {int(k.get('grid') or 0)} x {int(k.get('block') or 0)} | {int(k.get('regs') or 0)} | "
                    f"{(k.get('time_us') or 0):.1f} | {(k.get('dram_gbs') or 0):.0f} | {hcell} | {cell('alu')} | {cell('fma')} | {cell('lsu')} | "
                    f"{cell('xu')} | {(k.get('issue_pct') or 0):.1f} | {(k.get('occupancy_pct') or 0):.1f} | {', '.join(k['top_stalls'])} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
