#!/bin/bash
# round 2, visit 12: hybrid with bit output as the default -- host staging speed alone, decode tests, bench line
O=gpurun_out; mkdir -p $O
timeout 300 python tools/hostpack_bench.py 512 > $O/hostpack_bench.log 2>&1; cat $O/hostpack_bench.log
( timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_threads.py -m gpu -x -q > $O/pytest_gpu_exp12.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_exp12.log ); tail -4 $O/pytest_gpu_exp12.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-methods > $O/bench_exp12.json 2> $O/bench_exp12.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_exp12.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['host_path'])
print('ceiling', d['e2e']['copy_ceiling']['info_gbps_if_arrays_are_copied_as_they_are'], 'variants', {k:v.get('value') for k,v in d['e2e']['variants'].items()})
print('bytes', d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'])
PY
