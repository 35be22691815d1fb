#!/usr/bin/env python3
"""FER / BER waterfall of the engine's Monte-Carlo loop (ldpc_b200_simulate: fused producer + decoder + counters) for the
BASELINE configurations, one B200.  Per point: rounds of 2048 groups until >= `--errors` frame errors or `--max-frames`.

    python tools/waterfall.py [--out profiles/r01_waterfall] [--errors 300] [--max-frames 3.4e7] [--configs 1,2]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/waterfall.py ...   (frames sharded over the GPUs)

Under torchrun every rank simulates its own rounds (global frame index = Philox subsequence, so the frames are the ones a
single GPU would have drawn) and the counters are summed with the library's own NCCL all-reduce
(ldpc_b200_comm_init / ldpc_b200_allreduce_counters; main.cpp:170-182) after every round; all ranks evaluate the stop rule.

Golden codeword (FAKE_ENCODE, CSimulate.cpp:3), Philox seed 101 (CSimulate.cpp:11 uses 101 for thread 0), MaxIteration 15
for the early-stopping decoders (what the reference ships in Profile.txt) and 6 for NMS.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests"):
    sys.path.insert(0, str(ROOT / p))
import numpy as np  # noqa: E402

import ldpc_b200  # noqa: E402
import llrgen  # noqa: E402

K = 14592
CONFIGS = [
    # name, method, lut, scale, mod, interleave, max_iter, Eb/N0 grid
    ("NMS 26/26, QPSK, 6 it", 0, -1, 13.0, 2, 1, 6, np.arange(3.0, 4.61, 0.1)),
    ("FAID3 + DTBF, QPSK, 15 it", 2, 0, 13.0, 2, 1, 15, np.arange(3.0, 4.21, 0.1)),
    ("hybrid FAID + 2B1C, QPSK, scale 12.5, 15 it", 5, 3, 12.5, 2, 1, 15, np.arange(3.0, 4.21, 0.1)),
    ("OMS 1/6 + DTBF, 16-QAM I=4, 15 it", 4, -1, 13.0, 4, 4, 15, np.arange(6.8, 8.41, 0.2)),
    ("OMS 1/6 + DTBF, 64-QAM I=6, 15 it", 4, -1, 13.0, 6, 6, 15, np.arange(11.0, 12.81, 0.2)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(ROOT / "profiles" / "r01_waterfall"))
    ap.add_argument("--errors", type=int, default=300)
    ap.add_argument("--max-frames", type=float, default=3.4e7)
    ap.add_argument("--round-groups", type=int, default=2048)
    ap.add_argument("--configs", default="", help="comma-separated indices into CONFIGS (default: all)")
    ap.add_argument("--ebn0", default="", help="comma-separated Eb/N0 points overriding the grids")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # only to hand the NCCL unique id to the other ranks
        box = [ldpc_b200.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    cw = llrgen.golden_codeword()
    res = []
    sel = [int(x) for x in args.configs.split(",")] if args.configs else range(len(CONFIGS))
    for name, method, lut, scale, mod, il, mi, grid in (CONFIGS[i] for i in sel):
        if args.ebn0:
            grid = [float(x) for x in args.ebn0.split(",")]
        cfg = ldpc_b200.default_config(method, lut)
        cfg.scale, cfg.mod_type, cfg.interleave_mod_type, cfg.max_iteration = scale, mod, il, mi
        cfg.chunk_groups = args.round_groups
        cfg.device = local
        pts = []
        with ldpc_b200.Decoder(cfg) as dec:
            if world > 1:
                dec.comm_init(uid, rank, world)
            for eb in grid:
                cnt = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
                rnd = 0
                t0 = time.perf_counter()
                while cnt[1] < args.errors and cnt[0] < args.max_frames:
                    first = (rnd * world + rank) * args.round_groups * 32
                    c = dec.simulate(float(eb), 101, first, args.round_groups, codeword=cw)
                    if world > 1:
                        c = dec.allreduce_counters(c)
                    cnt += c
                    rnd += 1
                dt = time.perf_counter() - t0
                fr, fe, be, groups, its = (int(cnt[i]) for i in (0, 1, 2, 4, 5))
                pts.append({"ebn0_db": round(float(eb), 2), "frames": fr, "frame_errors": fe, "bit_errors": be, "fer": fe / fr, "ber": be / (fr * K),
                            "avg_min_sum_iterations": its / max(1, groups), "seconds": dt, "info_gbps": fr * K / dt / 1e9})
                if rank == 0:
                    print(f"{name}: {eb:.2f} dB  FER {fe / fr:.3e}  BER {be / (fr * K):.3e}  frames {fr}  its {its / max(1, groups):.2f}  {fr * K / dt / 1e9:.1f} Gbit/s", flush=True)
                if fe == 0:
                    break
        res.append({"config": name, "n_gpus": world, "method": method, "lut": lut, "scale": scale, "mod_type": mod, "interleave": il, "max_iteration": mi, "points": pts})
    if rank != 0:
        return
    Path(args.out + ".json").write_text(json.dumps(res, indent=1))
    with open(args.out + ".md", "w") as f:
        f.write(f"FER / BER of `ldpc_b200_simulate` on {world} B200 (golden codeword, Philox seed 101; `tools/waterfall.py`).  Each point stops at "
                f"{args.errors} frame errors or {args.max_frames:.1e} frames; throughput is whole-loop (producer + decoder + counters).\n\n")
        for r in res:
            f.write(f"**{r['config']}**\n\n| Eb/N0 dB | frames | FER | BER | avg min-sum iterations per group | Gbit/s |\n|---|---|---|---|---|---|\n")
            for p in r["points"]:
                f.write(f"| {p['ebn0_db']:.2f} | {p['frames']} | {p['fer']:.3e} | {p['ber']:.3e} | {p['avg_min_sum_iterations']:.2f} | {p['info_gbps']:.1f} |\n")
            f.write("\n")


if __name__ == "__main__":
    main()
