#!/bin/bash
# Round-2 GPU visit 9: FAID_M kinds with the (v + 31 | sign) word kept in a register (A/B on one box) + parity of that variant.
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp9.log; : > $L
for rep in 1 2; do
  timeout 300 python tools/nms_ab.py 2,5 1024 3.6 >> $L 2>&1
  LDPC_B200_LIB=$PWD/build/variants/faidm_reg.so timeout 300 python tools/nms_ab.py 2,5 1024 3.6 >> $L 2>&1
done
timeout 300 python tools/nms_ab.py 2,5 1024 4.2 >> $L 2>&1
LDPC_B200_LIB=$PWD/build/variants/faidm_reg.so timeout 300 python tools/nms_ab.py 2,5 1024 4.2 >> $L 2>&1
cat $L
( LDPC_B200_LIB=$PWD/build/variants/faidm_reg.so timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_at_scale.py -m gpu -q -k "2 or 5 or faid or FAID or hybrid" > $O/pytest_gpu_r02i_faidm_reg.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02i_faidm_reg.log )
tail -5 $O/pytest_gpu_r02i_faidm_reg.log
