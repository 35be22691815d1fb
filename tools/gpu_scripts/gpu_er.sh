#!/bin/bash
# Full GPU parity suite + timing of the erasure-mode kind next to the ordinary FAID kinds.
O=gpurun_out; TAG=${1:-er}
( timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -4 $O/pytest_gpu_$TAG.log
timeout 300 python tools/nms_ab.py 2 1024 3.6 > $O/quick_er_$TAG.log 2>&1; tail -8 $O/quick_er_$TAG.log
