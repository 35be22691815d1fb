"""The reference's threading model on the new boundary: ONE process, one object set per host thread
(CSimulate.cpp:218-278, main.cpp:166-172).  Several host threads create their own handles at the same moment (first use of
every per-device one-time initialisation in the library: code tables, kernel attributes, dlopen of NCCL) and decode /
simulate concurrently -- on one GPU, or spread over all visible GPUs.  Every thread's output must equal the oracle's.
Runs in a fresh interpreter so that the one-time initialisations really are raced."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import json, os, sys, threading
root = %(root)r
for p in ("mod-interleaveavx_multithreads-faid_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(root, p))
import numpy as np
import ldpc_b200, llrgen, pyoracle
ndev = %(ndev)d
methods = [0, 0, 2, 2, 4, 5, 1, 3]
T = len(methods)
fix = [np.concatenate([llrgen.qpsk_llr_groups(1, eb, scale=12.5 if m == 5 else 13.0, seed=500 + 10 * i + k)[0] for k, eb in enumerate((3.3, 3.7, 4.1))])
       for i, m in enumerate(methods)]
cw = llrgen.golden_codeword()
start = threading.Barrier(T)
res, errs = [None] * T, []

def work(i):
    try:
        cfg = ldpc_b200.default_config(methods[i], -1)
        cfg.device = i %% ndev
        cfg.chunk_groups, cfg.n_streams = 2, 2
        start.wait()
        with ldpc_b200.Decoder(cfg) as dec:                    # concurrent create()
            outs = []
            for rep in range(3):
                out, info = dec.decode(fix[i], want_info=True)  # concurrent launches, staging pools, finalize
                outs.append((out.copy(), [int(x) for x in info["bf_iters"]], [int(x) for x in info["its_per_group"]]))
            c1 = dec.simulate(3.7, 11 + i, 0, 4, codeword=cw).copy()
            c2 = dec.simulate(3.7, 11 + i, 0, 4).copy()         # encoder path (its own attribute)
            if i == 0:                                          # single-rank communicator from a worker thread
                dec.comm_init(ldpc_b200.nccl_unique_id(), 0, 1)
                c3 = dec.allreduce_counters(c1.copy())
                assert (c3 == c1).all()
        res[i] = (outs, c1, c2)
    except Exception as e:  # noqa
        errs.append(f"thread {i}: {e!r}")

ths = [threading.Thread(target=work, args=(i,)) for i in range(T)]
[t.start() for t in ths]
[t.join() for t in ths]
assert not errs, errs
orc = pyoracle.Oracle()
for i, m in enumerate(methods):
    ref, infos = orc.decode(orc.default_config(m, -1), fix[i])
    for out, bf, its in res[i][0]:
        assert (out == ref).all(), f"thread {i} method {m}: decoded bits differ from the oracle"
        assert bf == [x.bf_iters for x in infos] and its == [x.iters_executed for x in infos]
    assert res[i][1][0] == 128 and res[i][2][0] == 128
print("THREADS_OK " + json.dumps({"threads": T, "devices": ndev}))
"""


def test_one_handle_per_host_thread_concurrently(engine_lib, tmp_path):
    import torch
    ndev = max(1, torch.cuda.device_count())
    script = tmp_path / "threads_worker.py"
    script.write_text(WORKER % {"root": str(ROOT), "ndev": ndev})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("THREADS_OK ")][-1]
    assert json.loads(line[len("THREADS_OK "):])["threads"] == 8


def test_single_rank_communicator_and_teardown(engine_lib):
    """The library's own NCCL path (ldpc_b200_nccl_unique_id / comm_init / allreduce_counters, ncclCommDestroy in destroy())
    on whatever the lease offers: with one rank the all-reduce is the identity.  Twice, to cover communicator teardown."""
    import ldpc_b200
    for rep in range(2):
        with ldpc_b200.Decoder(ldpc_b200.default_config(0, -1)) as dec:
            dec.comm_init(ldpc_b200.nccl_unique_id(), 0, 1)
            c = np.arange(ldpc_b200.NUM_COUNTERS, dtype=np.uint64) * 3 + rep
            got = dec.allreduce_counters(c.copy())
            assert (got == c).all()
            # counters of a real round go through unchanged as well
            import llrgen
            r = dec.simulate(3.6, 5, 0, 2, codeword=llrgen.golden_codeword())
            assert (dec.allreduce_counters(r.copy()) == r).all()
    with pytest.raises(Exception):
        with ldpc_b200.Decoder(ldpc_b200.default_config(0, -1)) as dec:
            dec.allreduce_counters(np.zeros(ldpc_b200.NUM_COUNTERS, dtype=np.uint64))  # no communicator yet
