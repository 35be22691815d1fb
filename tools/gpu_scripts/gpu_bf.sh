#!/bin/bash
O=gpurun_out; TAG=${1:-bf}
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
for nf in 0 1; do
  if [ $nf = 1 ]; then export LDPC_B200_NO_FAST_BF=1; else unset LDPC_B200_NO_FAST_BF; fi
  for eb in 3.0 3.6; do echo "== no_fast_bf=$nf @ $eb dB"; timeout 300 python tools/nms_ab.py 2,3,4,5 1024 $eb; done
done 2>&1 | tee $O/bf_$TAG.log
