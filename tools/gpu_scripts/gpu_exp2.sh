#!/bin/bash
# Round-2 GPU visit 2: persistent decode kernel (parity subset + A/B) and the host-buffer path experiments.
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_at_scale.py tests/test_gpu_frames.py tests/test_gpu_random_configs.py -m gpu -q -x > $O/pytest_gpu_r02c.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02c.log )
tail -6 $O/pytest_gpu_r02c.log
L=$O/nms_ab_exp2.log; : > $L
timeout 300 python tools/nms_ab.py 0,1,2,4,5 1024 3.6 >> $L 2>&1
LDPC_B200_NO_SKEW=1 timeout 300 python tools/nms_ab.py 0,1,2 1024 3.6 >> $L 2>&1
for v in p2alu absfp16; do
  LDPC_B200_LIB=$PWD/build/variants/$v.so timeout 200 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1
done
timeout 200 python tools/nms_ab.py 0,2 1024 4.2 >> $L 2>&1
cat $L
timeout 120 python tools/copy_probe.py > $O/copy_probe_1gpu.json 2> $O/copy_probe_1gpu.err; cat $O/copy_probe_1gpu.json
timeout 900 python tools/e2e_exp.py 1024 > $O/e2e_exp2.log 2>&1; cat $O/e2e_exp2.log
