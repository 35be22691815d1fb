#!/bin/bash
# Round-2 GPU visit 15: message-word home / prefetch as template flags of the layer function (default) against pointer tests (cvdyn.so)
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp15.log; : > $L
for rep in 1 2; do
  timeout 300 python tools/nms_ab.py 0,1,2,4,5 1024 3.6 >> $L 2>&1
  LDPC_B200_LIB=$PWD/build/variants/cvdyn.so timeout 300 python tools/nms_ab.py 0,1,2,4,5 1024 3.6 >> $L 2>&1
done
timeout 300 python tools/nms_ab.py 1,2,5 1024 4.2 >> $L 2>&1
LDPC_B200_LIB=$PWD/build/variants/cvdyn.so timeout 300 python tools/nms_ab.py 1,2,5 1024 4.2 >> $L 2>&1
cat $L
( timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_at_scale.py -m gpu -q -x > $O/pytest_gpu_exp15.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_exp15.log )
tail -4 $O/pytest_gpu_exp15.log
