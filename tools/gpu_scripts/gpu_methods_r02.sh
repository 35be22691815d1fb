#!/bin/bash
# Round-2: the other BASELINE configurations as headline lines (bench.py --method 2 / 5 / 4) with their reference arms, and the
# Monte-Carlo waterfalls of the final build.
O=gpurun_out; mkdir -p $O
for m in 2 5 4; do
  timeout 600 python bench.py --method $m --steps 10 --warmup 3 --no-methods > $O/bench_r02_method$m.json 2> $O/bench_r02_method$m.err; echo "bench method $m rc=$?"
  timeout 200 python bench.py --impl reference --method $m --steps 2 --warmup 0 > $O/bench_r02_ref_method$m.json 2>> $O/bench_r02_method$m.err; echo "ref arm method $m rc=$?"
done
timeout 900 python tools/waterfall.py --out $O/r02_waterfall > $O/waterfall_r02.log 2>&1; echo "waterfall rc=$?"; tail -3 $O/waterfall_r02.log
python - <<'PY'
import json
for m in (2,5,4):
    d=json.load(open(f'gpurun_out/bench_r02_method{m}.json')); r=json.load(open(f'gpurun_out/bench_r02_ref_method{m}.json'))
    print(m, 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'frac', round(d['roofline']['frac'],3), 'ref', round(r['value'],3), 'kernel', d['kernel_ms_per_step'])
PY
