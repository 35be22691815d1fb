#!/bin/bash
# Round-2 GPU visit 10: hybrid host path vs number of staged / direct slots; tests of the current tree (FAID_M register form default).
O=gpurun_out; mkdir -p $O
timeout 900 python tools/e2e_hybrid.py 2048 > $O/e2e_hybrid_sweep.log 2>&1; cat $O/e2e_hybrid_sweep.log
( timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_at_scale.py tests/test_gpu_bounds_debug.py -m gpu -q > $O/pytest_gpu_r02j.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02j.log )
tail -5 $O/pytest_gpu_r02j.log
