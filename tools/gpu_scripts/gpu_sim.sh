#!/bin/bash
# fused vs separate producer: parity tests + throughput of ldpc_b200_simulate
O=gpurun_out; TAG=${1:-sim}
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
python - <<'PY' 2>&1 | tee $O/sim_$TAG.log
import sys, time, os
sys.path[:0] = ["mod-interleaveavx_multithreads-faid_b200", "tests"]
import numpy as np, ldpc_b200, llrgen
cw = llrgen.golden_codeword(); K = 14592; G = 1024
for method in (0, 1, 2, 5):
    for nofuse in (0, 1):
        if nofuse: os.environ["LDPC_B200_NO_FUSED_PRODUCER"] = "1"
        else: os.environ.pop("LDPC_B200_NO_FUSED_PRODUCER", None)
        cfg = ldpc_b200.default_config(method, -1); cfg.chunk_groups = 512
        with ldpc_b200.Decoder(cfg) as dec:
            for _ in range(2): dec.simulate(3.6, 101, 0, G, codeword=cw)
            t0 = time.time(); R = 5
            for i in range(R): c = dec.simulate(3.6, 101, i * G * 32, G, codeword=cw)
            dt = (time.time() - t0) / R
            t0 = time.time()
            for i in range(R): c2 = dec.simulate(3.6, 101, i * G * 32, G)
            dt2 = (time.time() - t0) / R
        print(f"method {method} fused={1-nofuse}: fixed codeword {G*32*K/dt/1e9:.2f} Gbit/s ({dt*1e3:.2f} ms/round), random info {G*32*K/dt2/1e9:.2f} Gbit/s, FER {c[1]/c[0]:.4f}")
PY
