#!/usr/bin/env python3
"""profiles/roofline_inputs.json entry of one kernel kind from an `ncu --set full` report of the CURRENT build:
DRAM bytes per frame (dram__bytes_read.sum + dram__bytes_write.sum of the kernel's launch / frames per launch) and the hash
of the kernel's static instruction mix (lib/sass_mix.json) -- bench.py reports `roofline.traffic` only while that hash matches
the library it loaded.
Usage: tools/update_roofline_inputs.py gpurun_out/x.ncu-rep NMS 32768 "how it was captured" """
import csv
import hashlib
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    rep, kind, frames, how = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))

    def bytes_of(key):
        v = float(d[key].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]

    rd, wr = bytes_of("dram__bytes_read.sum"), bytes_of("dram__bytes_write.sum")
    mix = json.loads((ROOT / "mod-interleaveavx_multithreads-faid_b200" / "lib" / "sass_mix.json").read_text())["kinds"][kind]
    p = ROOT / "profiles" / "roofline_inputs.json"
    allk = json.loads(p.read_text()) if p.exists() else {}
    allk[kind] = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_read_bytes_per_launch": rd, "dram_write_bytes_per_launch": wr,
                  "frames_per_launch": frames, "kernel": d.get("Kernel Name"), "duration_us": d.get("gpu__time_duration.sum"),
                  "kernel_sass_sha256": hashlib.sha256(json.dumps(mix, sort_keys=True).encode()).hexdigest(), "source": how}
    p.write_text(json.dumps(allk, indent=1) + "\n")
    print(json.dumps(allk[kind], indent=1))


if __name__ == "__main__":
    main()
