#!/bin/bash
# Round-2 GPU visit 18: ptxas -regUsageLevel 0 / 7 against the default 5
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp18.log; : > $L
for rep in 1 2; do
  timeout 300 python tools/nms_ab.py 0,1,2,5 1024 3.6 >> $L 2>&1
  for v in rul0 rul7; do LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/nms_ab.py 0,1,2,5 1024 3.6 >> $L 2>&1; done
done
cat $L
