#!/bin/bash
# Round-2 GPU visit 3: persistent vs one-CTA-per-two-pairs on the SAME box, with clocks and stall sampling.
O=gpurun_out; mkdir -p $O
L=$O/nms_ab_exp3.log; : > $L
smi() { nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader >> $L; }
smi
for rep in 1 2; do
  timeout 300 python tools/nms_ab.py 0,1,5 1024 3.6 >> $L 2>&1; smi
  LDPC_B200_NO_SKEW=1 timeout 300 python tools/nms_ab.py 0,1,5 1024 3.6 >> $L 2>&1
  LDPC_B200_LIB=$PWD/build/variants/nonpersistent.so timeout 300 python tools/nms_ab.py 0,1,5 1024 3.6 >> $L 2>&1; smi
done
cat $L
for v in default nonpersistent; do
  if [ "$v" = default ]; then unset LDPC_B200_LIB; else export LDPC_B200_LIB=$PWD/build/variants/$v.so; fi
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pair -c 1 -f -o $O/nms_r02_exp3_$v python tools/nms_ab.py 0 1024 3.6 > $O/ncu_exp3_$v.log 2>&1; echo "ncu $v rc=$?"
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pair -c 1 -f -o $O/oms_r02_exp3_$v python tools/nms_ab.py 1 1024 3.6 > $O/ncu_exp3_oms_$v.log 2>&1; echo "ncu oms $v rc=$?"
done
unset LDPC_B200_LIB
ls -la $O/*exp3*.ncu-rep
