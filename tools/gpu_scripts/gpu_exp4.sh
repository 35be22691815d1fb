#!/bin/bash
# Round-2 GPU visit 4: full GPU suite on the current tree, finalize (incremental syndrome) A/B, load-phase cost, hybrid host path, bench.
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02d.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02d.log )
tail -12 $O/pytest_gpu_r02d.log
L=$O/nms_ab_exp4.log; : > $L
timeout 300 python tools/nms_ab.py 0,1,2,3,4,5 1024 3.6 >> $L 2>&1
LDPC_B200_EXP_NOLOAD=1 timeout 200 python tools/nms_ab.py 0 1024 3.6 >> $L 2>&1
timeout 200 python tools/nms_ab.py 0 64 3.6 >> $L 2>&1
timeout 200 python tools/nms_ab.py 0 2048 3.6 >> $L 2>&1
cat $L
timeout 600 python tools/e2e_exp.py 2048 quick > $O/e2e_exp4.log 2>&1; grep -v "^Model\|^CPU\|^Thread\|^Core\|^Socket\|^L3\|^NUMA\|GPU0\|SYS\|NODE" $O/e2e_exp4.log
timeout 600 python bench.py > $O/bench_r02_exp4.json 2> $O/bench_r02_exp4.err; echo "bench rc=$?"; tail -3 $O/bench_r02_exp4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_exp4.json'))
for k in ('value','ms_per_step','gpu_launches','fer_at_3p6dB'): print(k, d[k])
print('e2e', d['e2e']['value'], d['e2e']['host_path'], d['e2e']['copy_ceiling'])
print('variants', {k:v.get('value') for k,v in d['e2e']['variants'].items()})
print('roofline', {k:d['roofline'][k] for k in ('achieved','peak','frac','traffic','traffic_source')})
print('packed', d['e2e_packed_layouts'].get('value'), 'sim', d['e2e_simulate_round'].get('value'))
print('other', {k:(round(v['value'],1), round(v['decode_ms'],2), round(v['finalize_ms'],2), round(v['roofline']['frac'],3)) for k,v in d['other_methods'].items()})
print('clocks', d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
