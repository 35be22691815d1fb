#!/usr/bin/env python3
"""Host <-> device copy ceiling of the box, with no kernels: every rank copies pinned host memory to its GPU and back,
both directions at once, all ranks at the same time.  This is what bounds the host-buffer call ldpc_b200_decode().

    python tools/copy_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/copy_probe.py

Prints one JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both together, the frames/s and
decoded-information Gbit/s those figures allow for the reference layouts (17 664 B in + 17 664 B out per frame) and for
the packed layouts (8 832 B in + 2 208 B out), plus the NUMA placement of each rank."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

N, K = 17664, 14592


def numa_info(local):
    info = {"cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        bdf = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
    except Exception:
        bdf = None
    try:
        import subprocess
        q = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        bdf = q.lower()
        if bdf.startswith("00000000:"):
            bdf = "0000:" + bdf[9:]
        p = f"/sys/bus/pci/devices/{bdf}/numa_node"
        info["gpu_bdf"] = bdf
        info["gpu_numa_node"] = int(open(p).read()) if os.path.exists(p) else None
    except Exception as e:  # pragma: no cover
        info["gpu_numa_node"] = None
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        info["numa_nodes"] = len(nodes)
    except Exception:
        info["numa_nodes"] = None
    return info


def timed(fn, seconds):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds:
        fn()
        torch.cuda.synchronize()
        n += 1
    return n, time.perf_counter() - t0


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    nb = 256 << 20
    h_a = torch.empty(nb, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(nb, dtype=torch.uint8).pin_memory()
    h_a.fill_(1); h_b.fill_(2)
    d_a = torch.empty(nb, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        h2d(); d2h()

    res = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("duplex", both)):
        for _ in range(2):
            fn()
        if world > 1:
            dist.barrier()
        n, dt = timed(fn, 1.0)
        res[name] = n * nb / dt / 1e9  # GB/s in each active direction
    mine = {"rank": rank, **res, **numa_info(local)}
    allr = [None] * world
    if world > 1:
        dist.all_gather_object(allr, mine)
    else:
        allr = [mine]
    if rank == 0:
        agg = {k: sum(r[k] for r in allr) for k in ("h2d", "d2h", "duplex")}
        out = {"n_gpus": world, "per_rank": allr, "aggregate_gbs": agg,
               "ceiling": {
                   "reference_layouts_frames_per_s": agg["duplex"] * 1e9 / N,
                   "reference_layouts_info_gbps": agg["duplex"] * 1e9 / N * K / 1e9,
                   "packed_layouts_info_gbps": min(agg["duplex"] * 1e9 / (N // 2), agg["duplex"] * 1e9 / (N // 8)) * K / 1e9,
                   "note": "duplex GB/s is per direction with both directions running; ldpc_b200_decode moves N bytes per frame each way"},
               "host": {"cpus": os.cpu_count()}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
