#!/usr/bin/env python3
"""Headline benchmark: decoded information Gbps of the layered QC-LDPC decoder on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--method M] [--groups G]

Workload (BASELINE.json configs[0] / SURVEY.md section 8d-1): 50G-PON (17664,14592) code, NMS
(DecodeMethod 0, Factor_1 = Factor_2 = 26, scale 13), MaxIteration 6, QPSK at Eb/N0 = 3.6 dB, synthetic frames of the
golden codeword produced on the device by the engine's own fused Philox producer.  One "step" = one pass of the
decoder over G groups of 32 frames per GPU (default 1024 groups = 579 MB of int8 LLRs in + 579 MB of decoded bits
out per GPU, both larger than the 126 MB L2).  For DecodeMethod 0 a step is ONE kernel launch: decode_pair_kernel writes
decodedBits itself.

  value  whole-job throughput with the LLRs resident in HBM when the timed region starts
  e2e    the same metric through the reference-facing C-ABI call ldpc_b200_decode() with HOST buffers
         (pinned): host->device and device->host copies are inside the timed region
  roofline       HBM view required by the bench contract (algorithmic bytes / kernel time vs measured copy peak)
  alu_roofline   the view that actually binds this kernel: integer lane-ops/s vs the measured issue-rate peak
  cpu_baseline   the reference's own AVX-512 decoder (oracle/_ref, compiled from /root/reference in the dev
                 container) or, if absent, the oracle port, timed on this box's host cores on a bounded sample

Under torchrun (N > 1) every rank decodes its own shard of groups (weak scaling, no data-path collective); the only
communication is one NCCL all-reduce of the 128 uint64 error/iteration counters per step, as in the reference's
join-and-sum (main.cpp:170-182).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "mod-interleaveavx_multithreads-faid_b200"))
sys.path.insert(0, str(ROOT / "tests"))

N, M, K = 17664, 3072, 14592
E = 70400
# Algorithmic HBM bytes per frame of the dominant kernel.  SURVEY 8d counts N int8 LLRs in + N/8 PACKED hard bits out + 8 B
# = 19,880 B; this bench writes the reference's decodedBits layout (one int8 per code bit, CLDPC.cpp:4796-4797) straight
# from the decode kernel, so the mandatory output is N bytes: 17,664 in + 17,664 out.
ALG_BYTES_PER_FRAME = 2 * 17664
ALG_BYTES_PER_FRAME_PACKED = 19880
NMS_LANE_OPS_PER_EDGE = 19.74    # the reference's own vector-ALU instruction count per edge update
METRIC = "decoded_info_gbps"
UNIT = "Gbit/s"
EBN0 = 3.6


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


ALU_PIPE_PEAK = 2.0  # warp-instructions / clk / SM: measured for LOP3, VIMNMX(.S16x2), VIADDMNMX, VABSDIFF4, SHF, PRMT
                     # (profiles/microbench/pipe_rates_r01.jsonl); these all share the one 64-lane "ALU" pipe
KIND_OF_METHOD = {0: "NMS", 1: "OMS", 2: "FAID_M", 3: "OMS", 4: "OMS", 5: "FAID_EF_M"}


def sass_mix():
    """Static per-edge instruction mix of the built kernels (tools/sass_mix.py, committed under profiles/)."""
    cands = sorted((ROOT / "profiles").glob("sass_mix_r*.json"))
    if not cands:
        return None, None
    return json.loads(cands[-1].read_text()), cands[-1].name


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (same grid)."""
    p = ROOT / "profiles" / "roofline_inputs.json"
    if p.exists():
        return json.loads(p.read_text())
    return {}


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
    sample = f"{threads} threads x >= {seconds:.0f} s looping Decode*() over 8 groups of 32 QPSK frames at {EBN0} dB, 6 iterations"
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
                        + f", MaxIteration=6, QPSK, Eb/N0={EBN0} dB, scale=13, golden codeword + Philox AWGN",
            "groups_per_gpu_per_step": args.groups, "frames_per_step": args.groups * 32 * n_gpus,
            "parallelism": f"groups sharded over {n_gpus} GPU(s), one counter all-reduce per step",
            "l2_policy": "inputs+outputs per step (1.16 GB/GPU at 1024 groups) exceed the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", type=int, default=0)
    ap.add_argument("--groups", type=int, default=1024, help="groups of 32 frames per GPU per step")
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
    if world > 1:
        # NCCL prints its version banner to STDOUT when the first communicator comes up; the contract is ONE JSON line on
        # stdout, so the banner is sent to stderr (fd-level redirect around init + the first collective)
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    G = args.groups
    cfg = ldpc_b200.default_config(args.method, -1)
    cfg.device = local
    # device-resident path: the whole step is ONE launch of the message-passing kernel + ONE finalize launch on one
    # stream, so the CUDA-event durations reported by the library are those of each kernel alone
    cfg.n_streams = 1
    cfg.chunk_groups = G
    dec = ldpc_b200.Decoder(cfg)
    # end-to-end path: chunks of 128 groups on 3 streams so that H2D, kernels and D2H of different chunks overlap
    cfg_e = ldpc_b200.default_config(args.method, -1)
    cfg_e.device = local
    cfg_e.n_streams = 3
    cfg_e.chunk_groups = min(128, G)
    dec_e = ldpc_b200.Decoder(cfg_e)

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
    d_info = torch.from_numpy(np.tile(cw[:K], 32).astype(np.int8)).cuda().repeat(G, 1)
    counters = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
    d_cnt = torch.zeros(ldpc_b200.NUM_COUNTERS, dtype=torch.int64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        dec.decode(d_fix, d_out)
        _, launches = dec.last_timing()
        ms = dec.last_timing_detail()
        if world > 1:  # the reference's only cross-worker exchange: sum the counters (main.cpp:170-182)
            d_cnt.zero_()
            d_cnt[0] = G * 32
            dist.all_reduce(d_cnt)
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

    # correctness guard inside the bench: the decoded frames are scored (not timed)
    c = dec.count_errors(d_info, d_out)
    fer = float(c[1]) / float(c[0])

    # ---- end to end through the C-ABI with host (pinned) buffers ----
    h_in = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
    h_out = ldpc_b200.PinnedArray((G, 32 * N), np.int8)
    h_in.array[:] = d_fix.cpu().numpy()
    for _ in range(2):
        dec_e.decode(h_in.array, h_out.array)
    barrier()
    e2e_steps = max(3, args.steps // 2)
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        dec_e.decode(h_in.array, h_out.array)
    barrier()
    e2e_elapsed = time.perf_counter() - t1
    e2e_ok = bool((h_out.array == d_out.cpu().numpy()).all())
    staging = dec_e.host_staging()

    # the same call with the other host-staging settings (informational; the headline e2e is the default handle above)
    e2e_variants = {}
    if not args.no_methods:
        for name, env in (("direct_copies", {"LDPC_B200_HOST_THREADS": "0"}),
                          ("bits_out", {"LDPC_B200_STAGE_OUT": "1", "LDPC_B200_STAGE_IN": "0"}),
                          ("nibbles_in_bits_out", {"LDPC_B200_STAGE_OUT": "1", "LDPC_B200_STAGE_IN": "1"})):
            saved = {k: os.environ.get(k) for k in ("LDPC_B200_HOST_THREADS", "LDPC_B200_STAGE_OUT", "LDPC_B200_STAGE_IN")}
            try:
                os.environ.update(env)
                with ldpc_b200.Decoder(cfg_e) as dv:
                    for _ in range(2):
                        dv.decode(h_in.array, h_out.array)
                    tv = time.perf_counter()
                    for _ in range(e2e_steps):
                        dv.decode(h_in.array, h_out.array)
                    dtv = time.perf_counter() - tv
                    stv = dv.host_staging()
                e2e_variants[name] = {"value": G * 32 * e2e_steps * K / dtv / 1e9, "unit": UNIT + " per GPU", "threads": stv["threads"],
                                      "h2d_bytes_per_step": stv["last_h2d_bytes"], "d2h_bytes_per_step": stv["last_d2h_bytes"],
                                      "ok": bool((h_out.array == d_out.cpu().numpy()).all())}
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
        hp_in = ldpc_b200.PinnedArray((G * 32, N // 2), np.uint8)
        hp_out = ldpc_b200.PinnedArray((G * 32, N // 32), np.uint32)
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
        e2e_packed = {"seconds": dt2, "h2d_bytes_per_step": G * 32 * N // 2, "d2h_bytes_per_step": G * 32 * (N // 32) * 4,
                      "bit_identical_to_device_path": ok, "call": "ldpc_b200_decode_packed (host pinned buffers)"}
        hp_in.free()
        hp_out.free()
    except Exception as ex:  # pragma: no cover
        e2e_packed = {"error": str(ex)}

    # ---- one whole Monte-Carlo round on the device (CSimulate::Run): producer + decoder + counters, counters only D2H ----
    e2e_sim = None
    try:
        for _ in range(2):
            dec.simulate(EBN0, 101, rank * G * 32, G, codeword=cw)
        barrier()
        t3 = time.perf_counter()
        sim_cnt = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
        for i in range(e2e_steps):
            dec.simulate(EBN0, 101, (rank + world * i) * G * 32, G, codeword=cw, counters=sim_cnt)
        barrier()
        e2e_sim = {"seconds": time.perf_counter() - t3, "fer": float(sim_cnt[1]) / float(max(1, sim_cnt[0])),
                   "call": "ldpc_b200_simulate (generate + decode + count on the device; 1 KB of counters D2H per step)"}
    except Exception as ex:  # pragma: no cover
        e2e_sim = {"error": str(ex)}

    # ---- the other DecodeMethods on the same resident frames (3.6 dB: groups run to MaxIteration) ----
    methods_ms = {}
    if not args.no_methods:
        for m in (1, 2, 3, 4, 5):
            try:
                cfg_m = ldpc_b200.default_config(m, -1)
                cfg_m.device = local
                cfg_m.n_streams = 1
                cfg_m.chunk_groups = G
                with ldpc_b200.Decoder(cfg_m) as dm:
                    for _ in range(2):
                        dm.decode(d_fix, d_out)
                    tot = 0.0
                    for _ in range(5):
                        dm.decode(d_fix, d_out)
                        tot += dm.last_timing()[0]
                    methods_ms[m] = tot / 5
            except Exception as ex:  # pragma: no cover
                methods_ms[m] = str(ex)

    # PCIe ceiling of this box for the e2e figure (not timed as part of any step): simultaneous pinned H2D + D2H
    pcie = {}
    try:
        nb = 256 << 20
        hp_a = torch.empty(nb, dtype=torch.uint8).pin_memory()
        hp_b = torch.empty(nb, dtype=torch.uint8).pin_memory()
        dv_a = torch.empty(nb, dtype=torch.uint8, device="cuda")
        dv_b = torch.empty(nb, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        tp = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s1):
                dv_a.copy_(hp_a, non_blocking=True)
            with torch.cuda.stream(s2):
                hp_b.copy_(dv_b, non_blocking=True)
        torch.cuda.synchronize()
        dtp = time.perf_counter() - tp
        pcie = {"duplex_gbs_each_way": 4 * nb / dtp / 1e9, "how": "4 x 256 MiB pinned H2D and D2H concurrently on two streams"}
        del hp_a, hp_b, dv_a, dv_b
    except Exception as ex:  # pragma: no cover
        pcie = {"error": str(ex)}

    t = torch.tensor([elapsed, e2e_elapsed, (e2e_packed or {}).get("seconds", 0.0), (e2e_sim or {}).get("seconds", 0.0)],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed, e2e_elapsed, packed_elapsed, sim_elapsed = [float(x) for x in t.cpu()]

    if rank == 0:
        frames_step = G * 32 * world
        value = frames_step * args.steps * K / elapsed / 1e9
        e2e_value = frames_step * e2e_steps * K / e2e_elapsed / 1e9
        hbm_peak, peak_src = measured_peaks()
        # dominant kernel = decode_pair_kernel (one launch per step on this handle)
        k_s = decode_ms / 1e3 / args.steps  # rank 0's own average launch duration, CUDA events on the launching stream
        frames_rank = G * 32
        ach_gbs = frames_rank * ALG_BYTES_PER_FRAME / k_s / 1e9
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        edge_updates = frames_rank * 6 * E / k_s
        mix, mix_src = sass_mix()
        kind = KIND_OF_METHOD.get(args.method, "NMS")
        alu_per_edge = (mix or {}).get(kind, {}).get("per_edge", {}).get("alu")
        all_per_edge = (mix or {}).get(kind, {}).get("per_edge_total")
        pair_edges_per_clk_sm = edge_updates / 2 / 32 / 148 / clk     # warp-level frame-pair edge updates per clk per SM
        traffic = ncu_traffic().get(kind)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": staging["last_h2d_bytes"], "d2h_bytes_per_step": staging["last_d2h_bytes"],
                    "steps": e2e_steps, "bit_identical_to_device_path": e2e_ok, "pcie": pcie,
                    "call": "ldpc_b200_decode (reference fixInput -> decodedBits int8 layouts, host buffers of G*32*N bytes each way)",
                    "host_staging": {"threads": staging["threads"], "llr_nibbles_in": staging["stage_in"], "decision_bits_out": staging["stage_out"],
                                     "note": "bytes_per_step are what crossed PCIe; decisions travel as bits and the library's host threads expand them into the caller's int8 decodedBits inside the timed region"},
                    "variants": e2e_variants},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "traffic_source": (traffic or {}).get("source"),
                         "peak_source": peak_src,
                         "note": "HBM is not the binding resource of this kernel (35,328 B/frame with int8 decodedBits, 19,880 B/frame with the packed output of ldpc_b200_decode_packed); see alu_roofline"},
            "alu_roofline": {"bound": "ALU pipe issue (integer min/max, LOP3, VABSDIFF4, SHF: 64 lanes/clk/SM)",
                             "achieved": None if alu_per_edge is None else alu_per_edge * pair_edges_per_clk_sm,
                             "peak": ALU_PIPE_PEAK, "unit": "ALU-pipe warp-inst/clk/SM",
                             "frac": None if alu_per_edge is None else alu_per_edge * pair_edges_per_clk_sm / ALU_PIPE_PEAK,
                             "alu_inst_per_pair_edge": alu_per_edge, "all_inst_per_pair_edge": all_per_edge,
                             "issue_slots_used_per_clk_sm": None if all_per_edge is None else all_per_edge * pair_edges_per_clk_sm,
                             "edge_updates_per_s": edge_updates, "instruction_mix_source": mix_src,
                             "note": "achieved = static ALU-pipe instructions per frame-pair edge update (SASS of the shipped kernel) x measured edge-update rate / sampled SM clock; cross-checked by ncu sm__inst_executed_pipe_alu in profiles/"},
            "kernel_ms_per_step": {"decode_pair_kernel": decode_ms / args.steps, "finalize_kernel": finalize_ms / args.steps},
            "fer_at_3p6dB": fer,
        }
        if e2e_packed and "seconds" in e2e_packed:
            e2e_packed["value"] = frames_step * e2e_steps * K / packed_elapsed / 1e9
            e2e_packed["unit"] = UNIT
        if e2e_sim and "seconds" in e2e_sim:
            e2e_sim["value"] = frames_step * e2e_steps * K / sim_elapsed / 1e9
            e2e_sim["unit"] = UNIT
        out["e2e_packed_layouts"] = e2e_packed
        out["e2e_simulate_round"] = e2e_sim
        out["other_methods"] = {
            str(m): ({"kernel_ms": v, "value": frames_rank * K / (v * 1e-3) / 1e9, "unit": UNIT + " per GPU, LLRs resident"}
                     if isinstance(v, float) else {"error": v}) for m, v in methods_ms.items()}
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
