#!/bin/bash
# Round-2 GPU visit 16: prefetch of the next layer's message words after phase 2 (prelate.so) against between the phases (default)
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp16.log; : > $L
for rep in 1 2 3; do
  timeout 300 python tools/nms_ab.py 0,1,4 1024 3.6 >> $L 2>&1
  LDPC_B200_LIB=$PWD/build/variants/prelate.so timeout 300 python tools/nms_ab.py 0,1,4 1024 3.6 >> $L 2>&1
done
cat $L
