#!/bin/bash
# Round-2, last visit: the headline line again (now with roofline.traffic matching the shipped kernel and the host-DRAM model of the
# e2e path), its reference arm, and the per-method lines with their reference arms, all on one box.
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r02_ref_final2.json 2> $O/bench_r02_final2.err; echo "bench ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r02_final2.json 2>> $O/bench_r02_final2.err; echo "bench rc=$?"; tail -2 $O/bench_r02_final2.err
for m in 2 5 4; do
  timeout 600 python bench.py --method $m --steps 10 --warmup 3 --no-methods > $O/bench_r02_method$m.json 2> $O/bench_r02_method$m.err; echo "bench method $m rc=$?"
  timeout 200 python bench.py --impl reference --method $m --steps 2 --warmup 0 > $O/bench_r02_ref_method$m.json 2>> $O/bench_r02_method$m.err; echo "ref arm method $m rc=$?"
done
( timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -k "multi_rank or hybrid or staging" > $O/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_final2.log ); tail -3 $O/pytest_gpu_final2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_final2.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/bench_r02_ref_final2.json').read().strip().splitlines()[-1])
print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ref', round(r['value'],2), 'traffic', d['roofline']['traffic'], 'frac', round(d['roofline']['frac'],3), 'issue', round(d['roofline']['issue_slot_frac'],3))
print('host dram', d['e2e']['host_dram_model'])
print('variants', {k:round(v.get('value',0),1) for k,v in d['e2e']['variants'].items()}, 'ceiling', d['e2e']['copy_ceiling']['info_gbps_if_arrays_are_copied_as_they_are'])
for m in (2,5,4):
    d=json.loads(open(f'gpurun_out/bench_r02_method{m}.json').read().strip().splitlines()[-1]); r=json.loads(open(f'gpurun_out/bench_r02_ref_method{m}.json').read().strip().splitlines()[-1])
    print(m, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ceiling', round(d['e2e']['copy_ceiling']['info_gbps_if_arrays_are_copied_as_they_are'],1), 'frac', round(d['roofline']['frac'],3), 'ref', round(r['value'],3))
PY
