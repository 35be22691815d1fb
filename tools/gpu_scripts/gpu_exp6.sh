#!/bin/bash
# Round-2 GPU visit 6: BF stage with inert-frame compaction, nibble loader; full suite + per-method timing + bench.
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02f.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02f.log )
tail -8 $O/pytest_gpu_r02f.log
L=$O/nms_ab_exp6.log; : > $L
timeout 300 python tools/nms_ab.py 0,1,2,3,4,5 1024 3.6 >> $L 2>&1
timeout 300 python tools/nms_ab.py 1,2,3,4,5 1024 3.9 >> $L 2>&1
LDPC_B200_NO_FAST_BF=1 timeout 300 python tools/nms_ab.py 4 1024 3.6 >> $L 2>&1
cat $L
timeout 600 python bench.py > $O/bench_r02_exp6.json 2> $O/bench_r02_exp6.err; echo "bench rc=$?"; tail -3 $O/bench_r02_exp6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_exp6.json'))
for k in ('value','ms_per_step','gpu_launches','fer_at_3p6dB'): print(k, d[k])
print('e2e', d['e2e']['value'], d['e2e']['copy_ceiling']['info_gbps_if_arrays_are_copied_as_they_are'])
print('variants', {k:v.get('value') for k,v in d['e2e']['variants'].items()})
print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac','traffic')})
print('packed', d['e2e_packed_layouts'].get('value'), 'sim', d['e2e_simulate_round'].get('value'))
print('other', {k:(round(v['value'],1), round(v['decode_ms'],2), round(v['finalize_ms'],2), round(v['roofline']['frac'],3)) for k,v in d['other_methods'].items()})
print('clocks', d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
