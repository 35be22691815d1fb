"""Two ranks, two GPUs: frames sharded by global frame index, counters summed with the library's own NCCL all-reduce
(ldpc_b200_comm_init / ldpc_b200_allreduce_counters; the reference sums its per-thread counters in main.cpp:170-182).
The totals must equal what ONE GPU counts for the same global frames.  Skipped on single-GPU boxes (the gloo version of
the sharding logic runs on CPU in tests/test_multirank_cpu.py)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import llrgen

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import json, os, sys
sys.path.insert(0, os.path.join(%(root)r, "mod-interleaveavx_multithreads-faid_b200")); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
import ldpc_b200, llrgen
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo")
box = [ldpc_b200.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
cfg = ldpc_b200.default_config(2, 0); cfg.device = local; cfg.chunk_groups = 64
cw = llrgen.golden_codeword()
tot = np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64)
with ldpc_b200.Decoder(cfg) as dec:
    dec.comm_init(box[0], rank, world)
    for rnd in range(3):
        c = dec.simulate(3.5, 101, (rnd * world + rank) * 64 * 32, 64, codeword=cw)
        tot += dec.allreduce_counters(c)
if rank == 0:
    print("COUNTERS " + json.dumps([int(x) for x in tot]))
"""


def test_two_gpus_count_what_one_gpu_counts(engine_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import ldpc_b200
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": str(ROOT)})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("COUNTERS ")][-1]
    two = np.array(json.loads(line[len("COUNTERS "):]), dtype=np.uint64)
    cfg = ldpc_b200.default_config(2, 0)
    cfg.chunk_groups = 64
    with ldpc_b200.Decoder(cfg) as dec:
        one = dec.simulate(3.5, 101, 0, 6 * 64, codeword=llrgen.golden_codeword())
    assert one[0] == 6 * 64 * 32 and 0 < one[1] < one[0]
    assert (two == one).all(), (two[:8], one[:8])
