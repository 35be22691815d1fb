#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i`) into a small JSON/markdown pair under profiles/.
Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name"), "id": d.get("ID")}
        for key in KEYS:
            if key in d:
                k[key] = {"value": d[key], "unit": u[key]}
        stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")}
        k["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        res.append(k)
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        for k in res:
            f.write(f"## {k['kernel']} (launch {k['id']})\n\n| metric | value | unit |\n|---|---|---|\n")
            for key in KEYS:
                if key in k:
                    f.write(f"| {key} | {k[key]['value']} | {k[key]['unit']} |\n")
            f.write("\nTop stall reasons (warps stalled per issued instruction): "
                    + ", ".join(f"{a} {b:.2f}" for a, b in k["stall_warps_per_issue"].items()) + "\n\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
