#!/bin/bash
# A/B of kernel variants (build/variants/*.so) + optional ncu capture of the default library.
#   gpurun --timeout 1200 -- 'bash tools/gpu_ab.sh TAG [methods] [ncu]'
TAG=${1:-ab}
METHODS=${2:-0,1,2,5}
O=gpurun_out
mkdir -p $O
rm -f $O/variants_$TAG.log
( timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
for lib in default build/variants/*.so; do
  if [ "$lib" = default ]; then unset LDPC_B200_LIB; else export LDPC_B200_LIB=$PWD/$lib; fi
  echo "== $lib" >> $O/variants_$TAG.log
  timeout 300 python tools/quick_bench.py $METHODS 1024 3.6 >> $O/variants_$TAG.log 2>&1
done
unset LDPC_B200_LIB
cat $O/variants_$TAG.log
if [ "$3" = ncu ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pair -c 1 -f -o $O/prof_nms_$TAG python tools/prof_decode.py 0 1024 > $O/ncu_$TAG.log 2>&1
  tail -2 $O/ncu_$TAG.log
fi
