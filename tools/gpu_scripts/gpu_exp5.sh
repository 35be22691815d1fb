#!/bin/bash
# Round-2 GPU visit 5: new LLR loader, fused host staging pass, bounds-check build.
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_bounds_debug.py tests/test_gpu_parity_at_scale.py tests/test_gpu_threads.py -m gpu -q > $O/pytest_gpu_r02e.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02e.log )
tail -8 $O/pytest_gpu_r02e.log
L=$O/nms_ab_exp5.log; : > $L
timeout 300 python tools/nms_ab.py 0,1,2,4,5 1024 3.6 >> $L 2>&1
LDPC_B200_EXP_NOLOAD=1 timeout 200 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1
timeout 200 python tools/nms_ab.py 0 2048 3.6 >> $L 2>&1
cat $L
timeout 600 python tools/e2e_exp.py 2048 quick > $O/e2e_exp5.log 2>&1; grep "Gbit" $O/e2e_exp5.log
