#!/bin/bash
# Round-2: the default bench line under torchrun on N GPUs (final build), as the driver launches it
O=gpurun_out; mkdir -p $O
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_r02_${N}gpu_final.json 2> $O/bench_r02_${N}gpu_final.err; echo "bench rc=$?"; tail -2 $O/bench_r02_${N}gpu_final.err
python - "$N" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/bench_r02_%sgpu_final.json' % sys.argv[1]).read().strip().splitlines()[-1])
e=d['e2e']
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(e['value'],2), 'ceiling(equal shards)', round(e['copy_ceiling']['equal_shards_ceiling_info_gbps'],1), {k: e['host_path'][k] for k in ('threads','llr_nibbles_in','decision_bits_out')})
print('host dram', round(e['host_dram_model']['gbs_all_ranks'],1), 'variants', {k:round(v.get('value',0),1) for k,v in e['variants'].items()})
print('packed', round(d['e2e_packed_layouts']['value'],1), 'sim', round(d['e2e_simulate_round']['value'],1), 'collective', d['collective'].get('sum_of_frames_ok'))
PY
