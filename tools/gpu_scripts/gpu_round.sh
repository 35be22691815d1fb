#!/bin/bash
# One GPU-box visit: parity tests, microbenchmark, A/B of kernel variants, headline bench.  Usage (from the repo root):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$TAG.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
( cd profiles/microbench && timeout 120 ./pipe_rates > ../../$O/pipe_rates_$TAG.jsonl 2>&1 )
for v in default nibhi addrhi both; do
  if [ "$v" = default ]; then unset LDPC_B200_LIB; else export LDPC_B200_LIB=$PWD/build/variants/$v.so; fi
  [ "$v" != default ] && [ ! -f "$LDPC_B200_LIB" ] && continue
  echo "== $v" >> $O/variants_$TAG.log
  timeout 300 python tools/nms_ab.py 0,1,2,5 1024 3.6 >> $O/variants_$TAG.log 2>&1
done
unset LDPC_B200_LIB
cat $O/variants_$TAG.log
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
cat $O/bench_$TAG.json
