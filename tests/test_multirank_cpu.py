"""world_size-2 gloo test (CPU) of the multi-rank host logic: group sharding and the single counter all-reduce."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT / "mod-interleaveavx_multithreads-faid_b200"))
    from ldpc_b200 import sharding
    from ldpc_b200.abi import CNT_ERROR_FRAME, CNT_TEST_FRAME
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G = 50
    rounds = 0
    total = sharding.empty_counters()
    seen = []
    while not sharding.stop_rule(total):
        g0, f0 = sharding.shard(rank, world, G, rounds)
        seen.append((g0, f0))
        c = sharding.empty_counters()
        c[CNT_TEST_FRAME] = G * 32
        c[CNT_ERROR_FRAME] = 3 + rank  # pretend
        total = total + sharding.allreduce_counters(c, dist)
        rounds += 1
    out.put((rank, rounds, total[:2].tolist(), seen))
    dist.destroy_process_group()


def test_two_ranks_share_groups_and_stop_together():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, r0, t0, s0), (_, r1, t1, s1) = res
    assert r0 == r1 == 3          # 7 error frames per round over both ranks -> 21 >= 20 after 3 rounds
    assert t0 == t1 == [3 * 2 * 50 * 32, 21]
    groups = [g for g, _ in s0 + s1]
    assert len(set(groups)) == len(groups) and sorted(groups) == list(range(0, 300, 50))
    assert all(f == 32 * g for g, f in s0 + s1)
