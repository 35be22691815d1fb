#!/bin/bash
# tests on the default build, then A/B of build/variants at two Eb/N0 points
TAG=${1:-ab2}
O=gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
rm -f $O/variants_$TAG.log
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset LDPC_B200_LIB; else export LDPC_B200_LIB=$PWD/$lib; fi
  for eb in 3.0 3.6 4.2; do
    echo "== $lib @ $eb dB" >> $O/variants_$TAG.log
    timeout 300 python tools/nms_ab.py 0,1,2,5 1024 $eb >> $O/variants_$TAG.log 2>&1
  done
done
unset LDPC_B200_LIB
cat $O/variants_$TAG.log
