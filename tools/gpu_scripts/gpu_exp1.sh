#!/bin/bash
# Round-2 GPU visit 1: full GPU test-suite + NMS kernel A/B (pipe-balance variants, start-up skew of the second pair) + ncu.
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_exp1.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02b.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02b.log )
tail -15 $O/pytest_gpu_r02b.log
L=$O/nms_ab_exp1.log; : > $L
timeout 300 python tools/nms_ab.py 0 1024 3.6 0,200,400,600,800,1000,1300 >> $L 2>&1
for v in noscale16 p2alu absfp16 both; do
  LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 200 python tools/nms_ab.py 0 1024 3.6 0,700 >> $L 2>&1
done
timeout 300 python tools/nms_ab.py 1,2,4,5 1024 3.6 0,700 >> $L 2>&1
cat $L
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pair -c 1 -f -o $O/nms_r02_exp1 python tools/nms_ab.py 0 1024 3.6 > $O/ncu_exp1.log 2>&1; echo "ncu rc=$?"
ls -la $O/*.ncu-rep
