#!/usr/bin/env python3
"""Headline benchmark: decoded information Gbps of the layered QC-LDPC decoder on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--method M] [--groups G]

Workload (BASELINE.json configs[0] / SURVEY.md section 8d-1): 50G-PON (17664,14592) code, NMS
(DecodeMethod 0, Factor_1 = Factor_2 = 26, scale 13), MaxIteration 6, QPSK at Eb/N0 = 3.6 dB, synthetic frames of the
golden codeword produced on the device by the engine's own fused Philox producer.  One "step" = one pass of the decoder
over G groups of 32 frames per GPU (default 8192 groups = 262 144 frames = 4.6 GB of int8 LLRs in + 4.6 GB of decoded
bits out per GPU, far beyond the 126 MB L2; ~47 ms, so that the driver's 20 steps are a ~1 s timed region and the clock
sampler sees dozens of samples).  For DecodeMethod 0 a step is a handful of launches of ONE kernel (decode_pair_kernel
writes decodedBits itself), one per chunk of the handle's scratch.

  value     whole-job throughput with the LLRs resident in HBM when the timed region starts
  e2e       the same metric through the reference-facing C-ABI call ldpc_b200_decode() with HOST buffers (pinned): the
            host->device and device->host copies are inside the timed region
  roofline  the BINDING resource of the dominant kernel: the ALU pipe (integer min/max, LOP3, VABSDIFF4, SHF, PRMT;
            2 warp-instructions/clk/SM measured, profiles/microbench).  achieved = ALU-pipe instructions per frame-pair
            edge update (static count from the SASS of the very library this process loaded, written by build.py and
            checked by hash) x measured edge-update rate / sampled SM clock.  `traffic` = DRAM bytes per launch from the
            committed `ncu --set full` capture of the same kernel.
  hbm_roofline   the contract's HBM view (algorithmic bytes / kernel time vs the measured copy peak): not binding
  cpu_baseline   the reference's own AVX-512 decoder (oracle/_ref, compiled from /root/reference in the dev container)
                 or, if absent, the oracle port, timed on this box's host cores on a bounded sample

Under torchrun (N > 1) every rank decodes its own shard of groups (weak scaling, no data-path collective); the only
communication is ONE all-reduce of the 128 uint64 error / iteration counters per step, as in the reference's join-and-sum
(main.cpp:170-182) -- issued through the LIBRARY's own communicator (ldpc_b200_comm_init / ldpc_b200_allreduce_counters,
NCCL over NVLink); torch.distributed only distributes the NCCL id and takes the max of the timings.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "mod-interleaveavx_multithreads-faid_b200"
sys.path.insert(0, str(PKG))
sys.path.insert(0, str(ROOT / "tests"))

N, M, K = 17664, 3072, 14592
E = 70400
# Algorithmic HBM bytes per frame of the dominant kernel.  SURVEY 8d counts N int8 LLRs in + N/8 PACKED hard bits out + 8 B
# = 19,880 B; this bench writes the reference's decodedBits layout (one int8 per code bit, CLDPC.cpp:4796-4797) straight
# from the decode kernel, so the mandatory output is N bytes: 17,664 in + 17,664 out.
ALG_BYTES_PER_FRAME = 2 * 17664
NMS_LANE_OPS_PER_EDGE = 19.74    # the reference's own vector-ALU instruction count per edge update
METRIC = "decoded_info_gbps"
UNIT = "Gbit/s"
EBN0 = 3.6
MAX_ITER = 6
ALU_PIPE_PEAK = 2.0  # warp-instructions / clk / SM: measured for LOP3, VIMNMX(.S16x2), VIADDMNMX, VABSDIFF4, SHF, PRMT
                     # (profiles/microbench/pipe_rates_r01.jsonl); these all share the one 64-lane "ALU" pipe
KIND_OF_METHOD = {0: "NMS", 1: "OMS", 2: "FAID_M", 3: "OMS", 4: "OMS", 5: "FAID_EF_M"}
E2E_GROUPS_CAP = 2048  # groups per e2e step (2 x 1.16 GB of pinned host memory); more e2e steps make up the duration


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sha256_of(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def sass_mix(lib_path):
    """Static per-edge instruction mix of the kernels of the library THIS process loaded.  build.py writes lib/sass_mix.json
    next to the .so; if it is missing or was made from another binary it is regenerated here (cuobjdump), loudly."""
    lib_path = Path(lib_path)
    mix_path = lib_path.parent / "sass_mix.json"
    sha = sha256_of(lib_path)
    if mix_path.exists():
        d = json.loads(mix_path.read_text())
        if d.get("lib_sha256") == sha:
            return d["kinds"], f"{mix_path.name} (sha256 {sha[:12]} = loaded library)"
        sys.stderr.write(f"bench.py: {mix_path} does not belong to the loaded library; regenerating\n")
    sys.path.insert(0, str(ROOT / "tools"))
    import sass_mix as sm
    d = sm.mix_of(lib_path)
    try:
        mix_path.write_text(json.dumps(d, indent=1) + "\n")
    except OSError:
        pass
    return d["kinds"], f"regenerated from {lib_path.name} (sha256 {sha[:12]})"


def ncu_traffic(kind, lib_sha):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture; null (and said so) when the
    capture was made from a different build of the kernels than the one loaded now."""
    p = ROOT / "profiles" / "roofline_inputs.json"
    if not p.exists():
        return None, "no committed capture"
    d = json.loads(p.read_text()).get(kind)
    if not d:
        return None, "no committed capture for this kernel"
    note = d.get("source", "")
    if d.get("kernel_sass_sha256") and d.get("kernel_sass_sha256") != kernel_sass_sha(kind):
        return None, f"stale: {note} was captured from another build of this kernel"
    return d, note


_KSHA = {}


def kernel_sass_sha(kind):
    """Hash of the instruction mix entry of a kernel kind (changes whenever its SASS loop changes)."""
    return _KSHA.get(kind)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(method, seconds, threads=None):
    """Reference decoder on the host cores: all threads, private CLDPC each, decode-only, `seconds` of wall time."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import llrgen
    import pyoracle
    threads = threads or os.cpu_count() or 1
    groups, _ = llrgen.qpsk_llr_groups(8, EBN0, seed=3)
    sample = f"{threads} threads x >= {seconds:.0f} s looping Decode*() over 8 groups of 32 QPSK frames at {EBN0} dB, {MAX_ITER} iterations"
    if pyoracle.ref_available("faid3"):
        ref = pyoracle.Ref("faid3")
        cfg = pyoracle.Oracle().default_config(method)
        fps, done = ref.bench_decode(cfg, groups, threads, seconds)
        kind = "reference"
    else:
        orc = pyoracle.Oracle()
        cfg = orc.default_config(method)
        fps, done = orc.bench_decode(cfg, groups, threads, seconds)
        kind = "port"
    return {"value": fps * K / 1e9, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "frames": int(done)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(args.method, 1.0)
    per_step = 4.0
    cb = None
    for _ in range(args.steps):
        cb = cpu_baseline(args.method, per_step)
        vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    cb["sample"] = f"{args.steps} steps, each: " + cb["sample"].replace(">= 4 s", f">= {per_step:.0f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }))


def workload_config(args, n_gpus):
    return {"workload": f"50G-PON (17664,14592) QC-LDPC, DecodeMethod={args.method} "
                        + ("NMS Factor_1=Factor_2=26" if args.method == 0 else "reference shipped constants")
                        + f", MaxIteration={MAX_ITER}, QPSK, Eb/N0={EBN0} dB, scale=13, golden codeword + Philox AWGN",
            "groups_per_gpu_per_step": args.groups, "frames_per_step": args.groups * 32 * n_gpus,
            "parallelism": f"groups sharded over {n_gpus} GPU(s), one counter all-reduce per step",
            "l2_policy": f"inputs+outputs per step ({args.groups * 32 * 2 * N / 1e9:.2f} GB/GPU) exceed the 126 MB L2"}


def kernel_roofline(kind, mix, frames, decode_ms, clk_hz, sms=148):
    """Pipe roofline of decode_pair_kernel<kind> from its static ALU-pipe count and a measured launch duration."""
    m = (mix or {}).get(kind)
    if not m or decode_ms <= 0:
        return None
    edge_updates = frames * MAX_ITER * E / (decode_ms * 1e-3)
    pair_edges_per_clk_sm = edge_updates / 2 / 32 / sms / clk_hz
    alu = m["per_edge"]["alu"]
    return {"bound": "alu_pipe", "achieved": alu * pair_edges_per_clk_sm, "peak": ALU_PIPE_PEAK, "unit": "warp-inst/clk/SM",
            "frac": alu * pair_edges_per_clk_sm / ALU_PIPE_PEAK, "alu_inst_per_pair_edge": alu,
            "all_inst_per_pair_edge": m["per_edge_total"], "issue_slots_per_clk_sm": m["per_edge_total"] * pair_edges_per_clk_sm,
            "issue_slot_frac": m["per_edge_total"] * pair_edges_per_clk_sm / 4.0,  # 4 schedulers x 1 warp-instruction per clock
            "edge_updates_per_s": edge_updates}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", type=int, default=0)
    ap.add_argument("--groups", type=int, default=8192, help="groups of 32 frames per GPU per step")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-methods", action="store_true", help="skip the short per-DecodeMethod kernel timings")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import ldpc_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    lib_path = Path(os.environ.get("LDPC_B200_LIB", str(ldpc_b200.LIB_PATH)))
    mix, mix_src = sass_mix(lib_path) if rank == 0 else (None, None)
    if mix:
        for k, v in mix.items():
            _KSHA[k] = hashlib.sha256(json.dumps(v, sort_keys=True).encode()).hexdigest()

    G = args.groups
    cfg = ldpc_b200.default_config(args.method, -1)
    cfg.device = local
    # device-resident path: one stream, the largest chunks the handle's scratch allows; the CUDA-event durations reported by
    # the library are then those of each kernel alone
    cfg.n_streams = 1
    cfg.chunk_groups = min(G, 2048)
    dec = ldpc_b200.Decoder(cfg)

    if world > 1:
        # NCCL prints its version banner to STDOUT when the first communicator comes up; the contract is ONE JSON line on
        # stdout, so the banner is sent to stderr (fd-level redirect around both communicators' start-up)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            box = [ldpc_b200.nccl_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            dec.comm_init(box[0], rank, world)            # the product's own communicator (ncclCommInitRank via the C-ABI)
            warm = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
            warm[0] = 1
            assert int(dec.allreduce_counters(warm)[0]) == world
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # synthetic frames, generated on the device by the engine's fused producer; global frame index keeps the
    # stream independent of the GPU count
    import llrgen
    cw = llrgen.golden_codeword()
    tx_one = np.concatenate([np.tile(cw[:K], 32), np.tile(cw[K:], 32)]).astype(np.int8)
    d_tx = torch.from_numpy(tx_one).cuda().repeat(min(G, 64), 1)
    d_fix = torch.empty((G, 32 * N), dtype=torch.int8, device="cuda")
    for g0 in range(0, G, d_tx.shape[0]):
        g1 = min(G, g0 + d_tx.shape[0])
        d_fix[g0:g1] = dec.generate(d_tx[: g1 - g0], EBN0, 101, (rank * G + g0) * 32, g1 - g0)
    d_out = torch.empty_like(d_fix)
    step_counters = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
    reduced_frames = [0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        dec.decode(d_fix, d_out)
        _, launches = dec.last_timing()
        ms = dec.last_timing_detail()
        if world > 1:  # the reference's only cross-worker exchange: sum the counters (main.cpp:170-182)
            step_counters[:] = 0
            step_counters[0] = G * 32
            dec.allreduce_counters(step_counters)
            reduced_frames[0] = int(step_counters[0])
        return ms, launches

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t0 = time.perf_counter()
    decode_ms, finalize_ms, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        ms, nl = step_device()
        decode_ms += ms[0]
        finalize_ms += ms[1]
        launches += nl
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    comm_ok = world == 1 or reduced_frames[0] == G * 32 * world

    # correctness guard inside the bench: the decoded frames are scored (not timed), 1024 groups at a time
    tile = min(G, 1024)
    d_info = torch.from_numpy(np.tile(cw[:K], 32).astype(np.int8)).cuda().repeat(tile, 1)
    c = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
    for g0 in range(0, G, tile):
        g1 = min(G, g0 + tile)
        dec.count_errors(d_info[: g1 - g0], d_out[g0:g1], counters=c)
    fer = float(c[1]) / float(c[0])
    del d_info

    # ---- end to end through the C-ABI with host (pinned) buffers ----
    Ge = min(G, E2E_GROUPS_CAP)
    cfg_e = ldpc_b200.default_config(args.method, -1)
    cfg_e.device = local
    # the library's defaults for host arrays: 6 slots, chunks of 64 groups (staged) / 32 (copied as they are), so that H2D,
    # kernels, D2H and host staging of different chunks overlap
    dec_e = ldpc_b200.Decoder(cfg_e)
    h_in = ldpc_b200.PinnedArray((Ge, 32 * N), np.int8)
    h_out = ldpc_b200.PinnedArray((Ge, 32 * N), np.int8)
    h_in.array[:] = d_fix[:Ge].cpu().numpy()
    ref_out = d_out[:Ge].cpu().numpy()
    for _ in range(2):
        dec_e.decode(h_in.array, h_out.array)
    barrier()
    e2e_steps = max(3, int(round(args.steps * G / Ge / 2)))
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        dec_e.decode(h_in.array, h_out.array)
    barrier()
    e2e_elapsed = time.perf_counter() - t1
    e2e_ok = bool((h_out.array == ref_out).all())
    staging = dec_e.host_staging()
    routing = dec_e.last_routing()
    placement = dec_e.host_placement()

    # the same call with the other host-path settings (informational; the headline e2e is the default handle above)
    e2e_variants = {}
    if not args.no_methods:
        var_steps = max(3, e2e_steps // 4)
        for name, env in (("direct_copies", {"LDPC_B200_HOST_THREADS": "0"}),
                          ("staged_only", {"LDPC_B200_HYBRID": "0"}),
                          ("hybrid_direct_slot_bytes_out", {"LDPC_B200_HYBRID_OUT_BITS": "0"}),
                          ("bits_out", {"LDPC_B200_STAGE_OUT": "1", "LDPC_B200_STAGE_IN": "0"})):
            keys = ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN", "LDPC_B200_HYBRID", "LDPC_B200_HYBRID_OUT_BITS")
            saved = {k: os.environ.get(k) for k in keys}
            try:
                os.environ.update(env)
                with ldpc_b200.Decoder(cfg_e) as dv:
                    for _ in range(2):
                        dv.decode(h_in.array, h_out.array)
                    tv = time.perf_counter()
                    for _ in range(var_steps):
                        dv.decode(h_in.array, h_out.array)
                    dtv = time.perf_counter() - tv
                    stv = dv.host_staging()
                e2e_variants[name] = {"value": Ge * 32 * var_steps * K / dtv / 1e9, "unit": UNIT + " per GPU", "threads": stv["threads"],
                                      "h2d_bytes_per_step": stv["last_h2d_bytes"], "d2h_bytes_per_step": stv["last_d2h_bytes"],
                                      "ok": bool((h_out.array == ref_out).all())}
            except Exception as ex:  # pragma: no cover
                e2e_variants[name] = {"error": str(ex)}
            finally:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v

    # ---- the same frames through the engine's native packed layouts (nibble LLRs in, bit-packed decisions out) ----
    e2e_packed = None
    try:
        hp_in = ldpc_b200.PinnedArray((Ge * 32, N // 2), np.uint8)
        hp_out = ldpc_b200.PinnedArray((Ge * 32, N // 32), np.uint32)
        hp_in.array[:] = ldpc_b200.pack_llr(h_in.array)
        for _ in range(2):
            dec_e.decode_packed(hp_in.array, hp_out.array)
        barrier()
        t2 = time.perf_counter()
        for _ in range(e2e_steps):
            dec_e.decode_packed(hp_in.array, hp_out.array)
        barrier()
        dt2 = time.perf_counter() - t2
        ok = bool((ldpc_b200.unpack_hard(hp_out.array[: 64]).reshape(2, -1) == h_out.array[:2]).all())
        e2e_packed = {"seconds": dt2, "h2d_bytes_per_step": Ge * 32 * N // 2, "d2h_bytes_per_step": Ge * 32 * (N // 32) * 4,
                      "bit_identical_to_device_path": ok, "call": "ldpc_b200_decode_packed (host pinned buffers)"}
        hp_in.free()
        hp_out.free()
    except Exception as ex:  # pragma: no cover
        e2e_packed = {"error": str(ex)}

    # ---- one whole Monte-Carlo round on the device (CSimulate::Run): producer + decoder + counters, counters only D2H ----
    e2e_sim = None
    try:
        for _ in range(2):
            dec.simulate(EBN0, 101, rank * Ge * 32, Ge, codeword=cw)
        barrier()
        t3 = time.perf_counter()
        sim_cnt = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
        for i in range(e2e_steps):
            dec.simulate(EBN0, 101, (rank + world * i) * Ge * 32, Ge, codeword=cw, counters=sim_cnt)
        barrier()
        e2e_sim = {"seconds": time.perf_counter() - t3, "fer": float(sim_cnt[1]) / float(max(1, sim_cnt[0])),
                   "call": "ldpc_b200_simulate (generate + decode + count on the device; 1 KB of counters D2H per step)"}
    except Exception as ex:  # pragma: no cover
        e2e_sim = {"error": str(ex)}

    # ---- the other DecodeMethods on the same resident frames (3.6 dB: groups run to MaxIteration) ----
    methods_ms = {}
    Gm = min(G, 2048)
    if not args.no_methods:
        for m in (1, 2, 3, 4, 5):
            try:
                cfg_m = ldpc_b200.default_config(m, -1)
                cfg_m.device = local
                cfg_m.n_streams = 1
                cfg_m.chunk_groups = 1024
                with ldpc_b200.Decoder(cfg_m) as dm:
                    for _ in range(2):
                        dm.decode(d_fix[:Gm], d_out[:Gm])
                    tot = [0.0, 0.0]
                    for _ in range(5):
                        dm.decode(d_fix[:Gm], d_out[:Gm])
                        a, b = dm.last_timing_detail()
                        tot[0] += a / 5
                        tot[1] += b / 5
                    methods_ms[m] = tot
            except Exception as ex:  # pragma: no cover
                methods_ms[m] = str(ex)

    # Copy ceiling of this box for the e2e figure (not timed as part of any step): pinned H2D + D2H on two streams, ALL
    # ranks at the same time -- what ldpc_b200_decode can reach at most when the caller's arrays are copied as they are
    pcie = {}
    try:
        nb = 256 << 20
        hp_a = torch.empty(nb, dtype=torch.uint8).pin_memory()
        hp_b = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dv_a = torch.empty(nb, dtype=torch.uint8, device="cuda")
        dv_b = torch.empty(nb, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for _ in range(2):
            with torch.cuda.stream(s1):
                dv_a.copy_(hp_a, non_blocking=True)
            with torch.cuda.stream(s2):
                hp_b.copy_(dv_b, non_blocking=True)
        barrier()
        tp = time.perf_counter()
        for _ in range(8):
            with torch.cuda.stream(s1):
                dv_a.copy_(hp_a, non_blocking=True)
            with torch.cuda.stream(s2):
                hp_b.copy_(dv_b, non_blocking=True)
        torch.cuda.synchronize()
        dtp = time.perf_counter() - tp
        pcie = {"duplex_gbs_each_way": 8 * nb / dtp / 1e9, "how": "8 x 256 MiB pinned H2D and D2H concurrently on two streams, every rank at once"}
        del hp_a, hp_b, dv_a, dv_b
    except Exception as ex:  # pragma: no cover
        pcie = {"error": str(ex)}

    t = torch.tensor([elapsed, e2e_elapsed, (e2e_packed or {}).get("seconds", 0.0), (e2e_sim or {}).get("seconds", 0.0),
                      -pcie.get("duplex_gbs_each_way", 0.0)], dtype=torch.float64, device="cuda")
    tsum = torch.tensor([pcie.get("duplex_gbs_each_way", 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    elapsed, e2e_elapsed, packed_elapsed, sim_elapsed, neg_min_duplex = [float(x) for x in t.cpu()]
    duplex_sum = float(tsum.cpu()[0])

    if rank == 0:
        frames_step = G * 32 * world
        value = frames_step * args.steps * K / elapsed / 1e9
        e2e_value = Ge * 32 * world * e2e_steps * K / e2e_elapsed / 1e9
        hbm_peak, peak_src = measured_peaks()
        # dominant kernel = decode_pair_kernel; rank 0's own average duration per step, CUDA events on the launching stream
        k_ms = decode_ms / args.steps
        frames_rank = G * 32
        ach_gbs = frames_rank * ALG_BYTES_PER_FRAME / (k_ms * 1e-3) / 1e9
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        kind = KIND_OF_METHOD.get(args.method, "NMS")
        roof = kernel_roofline(kind, mix, frames_rank, k_ms, clk)
        if roof is None:
            raise SystemExit(f"bench.py: no instruction mix for kernel kind {kind} of {lib_path} (tools/sass_mix.py)")
        traffic, traffic_note = ncu_traffic(kind, None)
        launches_per_step = launches / args.steps
        roof.update({
            "traffic": None if not traffic else traffic.get("dram_bytes_per_frame", 0) * frames_rank / max(1.0, launches_per_step),
            "traffic_unit": "DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch)",
            "traffic_source": traffic_note, "instruction_mix_source": mix_src,
            "peak_source": "profiles/microbench/pipe_rates_r01.jsonl (LOP3 / VIMNMX / VIADDMNMX / VABSDIFF4 / SHF / PRMT: 2.0 each)",
            "algorithmic": {"lane_ops_per_edge_update": NMS_LANE_OPS_PER_EDGE,
                            "lane_ops_per_s": None if args.method != 0 else NMS_LANE_OPS_PER_EDGE * roof["edge_updates_per_s"],
                            "note": "SURVEY 8d: int8 lane-ops the reference's own row loop spends per edge update"},
            "note": "achieved = static ALU-pipe instructions per frame-pair edge update (SASS of the loaded library) x measured edge-update rate / sampled SM clock; ncu sm__inst_executed_pipe_alu of the same kernel is in profiles/"})
        ceiling_gbps = duplex_sum * 1e9 / N * K / 1e9 if duplex_sum > 0 else None
        # what the host's memory system moves per frame on this path (model from the bytes that crossed PCIe): the DMA itself, plus
        # read N + write N/2 for every nibble-packed frame and read N/8 + write N for every expanded one
        fr_e = Ge * 32
        b_in, b_out = staging["last_h2d_bytes"] / fr_e, staging["last_d2h_bytes"] / fr_e
        x_pack, y_exp = max(0.0, 2.0 * (1.0 - b_in / N)), max(0.0, (1.0 - b_out / N) / 0.875)
        dram_per_frame = b_in + b_out + 1.5 * N * x_pack + 1.125 * N * y_exp
        host_dram = {"bytes_per_frame": dram_per_frame, "packed_share": x_pack, "expanded_share": y_exp,
                     "gbs_all_ranks": dram_per_frame * e2e_value * 1e9 / K / 1e9,
                     "note": "model, not a counter: DMA reads / writes + the staging threads' reads and streaming writes; tools/hostpack_bench.py "
                             "measures 159-177 GB/s for a streaming copy / the fused staging pass on 16 threads of a 1-GPU box (profiles/r02_hostpack_bench.log)"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": staging["last_h2d_bytes"], "d2h_bytes_per_step": staging["last_d2h_bytes"],
                    "steps": e2e_steps, "groups_per_gpu_per_step": Ge, "bit_identical_to_device_path": e2e_ok,
                    "call": "ldpc_b200_decode (reference fixInput -> decodedBits int8 layouts, pinned host buffers of G*32*N bytes each way)",
                    "copy_ceiling": {"duplex_gbs_each_way_sum_over_ranks": duplex_sum, "min_over_ranks": -neg_min_duplex,
                                     "info_gbps_if_arrays_are_copied_as_they_are": ceiling_gbps,
                                     "e2e_over_ceiling": None if not ceiling_gbps else e2e_value / ceiling_gbps,
                                     "equal_shards_ceiling_info_gbps": None if neg_min_duplex >= 0 else -neg_min_duplex * world * 1e9 / N * K / 1e9,
                                     "e2e_over_equal_shards_ceiling": None if neg_min_duplex >= 0 else e2e_value / (-neg_min_duplex * world * 1e9 / N * K / 1e9),
                                     "note": "every rank moves the same number of frames, so the job ends with its slowest link: n_gpus x the minimum over ranks is what equal shards can reach",
                                     "how": pcie.get("how")},
                    "host_path": {"threads": staging["threads"], "llr_nibbles_in": staging["stage_in"], "decision_bits_out": staging["stage_out"],
                                  "staged_chunks": routing["staged_chunks"], "direct_chunks": routing["direct_chunks"],
                                  "numa_node": placement["numa_node"], "numa_cpus": placement["numa_cpus"],
                                  "note": "bytes_per_step are what crossed PCIe; staged chunks travel as nibbles / bits and the library's host threads pack / expand them inside the timed region, direct chunks are copied as they are"},
                    "host_dram_model": host_dram,
                    "variants": e2e_variants},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "hbm_roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "peak_source": peak_src,
                             "note": "not the binding resource: 35,328 B/frame with int8 decodedBits (19,880 B/frame with the packed output of ldpc_b200_decode_packed)"},
            "kernel_ms_per_step": {"decode_pair_kernel": decode_ms / args.steps, "finalize_kernel": finalize_ms / args.steps,
                                   "launches_per_step": launches_per_step},
            "collective": {"call": "ldpc_b200_allreduce_counters (ncclAllReduce of 128 uint64, the library's own communicator)",
                           "per_step": world > 1, "ranks": world, "sum_of_frames_ok": comm_ok},
            "fer_at_3p6dB": fer,
        }
        if e2e_packed and "seconds" in e2e_packed:
            e2e_packed["value"] = Ge * 32 * world * e2e_steps * K / packed_elapsed / 1e9
            e2e_packed["unit"] = UNIT
        if e2e_sim and "seconds" in e2e_sim:
            e2e_sim["value"] = Ge * 32 * world * e2e_steps * K / sim_elapsed / 1e9
            e2e_sim["unit"] = UNIT
        out["e2e_packed_layouts"] = e2e_packed
        out["e2e_simulate_round"] = e2e_sim
        other = {}
        for m, v in methods_ms.items():
            if not isinstance(v, list):
                other[str(m)] = {"error": v}
                continue
            r = kernel_roofline(KIND_OF_METHOD[m], mix, Gm * 32, v[0], clk)
            other[str(m)] = {"kernel_ms": v[0] + v[1], "decode_ms": v[0], "finalize_ms": v[1],
                             "value": Gm * 32 * K / ((v[0] + v[1]) * 1e-3) / 1e9, "unit": UNIT + " per GPU, LLRs resident", "groups": Gm,
                             "roofline": None if r is None else {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "alu_inst_per_pair_edge", "all_inst_per_pair_edge")}}
        out["other_methods"] = other
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.method, args.cpu_seconds)
        print(json.dumps(out))
    h_in.free()
    h_out.free()
    dec.close()
    dec_e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
