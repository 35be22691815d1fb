#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over small decodes of every kernel kind and the fused producer
O=gpurun_out; TAG=${1:-san}
cat > /tmp/san_work.py <<'PY'
import os, sys
sys.path[:0] = ["mod-interleaveavx_multithreads-faid_b200", "tests"]
import numpy as np, ldpc_b200, llrgen
fix, cw = llrgen.qpsk_llr_groups(2, 3.7, seed=3)
for m in (0, 1, 2, 3, 4, 5):
    cfg = ldpc_b200.default_config(m, -1)
    with ldpc_b200.Decoder(cfg) as dec:
        out = dec.decode(fix)
        c = dec.simulate(3.7, 5, 0, 2, codeword=cw)
        c2 = dec.simulate(3.7, 5, 64, 2)
    print(m, int(out.sum()), c[:4], c2[:4])
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_work.py > $O/sanitizer_${tool}_$TAG.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/sanitizer_${tool}_$TAG.log | tail -3
done
