#!/bin/bash
# Round-2 GPU visit 22: XOR median (ALU pipe, 2 instructions) for some of the triples of the two-smallest search instead of the fp16 one (FMA pipe, 4)
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp22.log; : > $L
for rep in 1 2 3; do
  timeout 120 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1
  for v in build/variants/x_*.so; do LDPC_B200_LIB=$PWD/$v timeout 120 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1; done
done
cat $L
