#!/bin/bash
# Round-2 GPU visit 21: schedule perturbations of the NMS kernel on top of the template flags (one box, 2 rounds)
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp21.log; : > $L
for rep in 1 2; do
  timeout 120 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1
  for v in build/variants/s_*.so; do LDPC_B200_LIB=$PWD/$v timeout 120 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1; done
done
cat $L
