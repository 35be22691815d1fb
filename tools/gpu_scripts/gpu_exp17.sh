#!/bin/bash
# Round-2 GPU visit 17: which of the two template flags (1 = message home, 2 = prefetch) each kernel family should take
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp17.log; : > $L
for rep in 1 2; do
  timeout 300 python tools/nms_ab.py 0,1,2,5 1024 3.6 >> $L 2>&1
  for v in nms_cv1 nms_cv2; do LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1; done
  for v in oms_cv1 oms_cv2; do LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/nms_ab.py 1 1024 3.6 >> $L 2>&1; done
  for v in faid_cv1 faid_cv2all; do LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/nms_ab.py 2,5 1024 3.6 >> $L 2>&1; done
done
cat $L
