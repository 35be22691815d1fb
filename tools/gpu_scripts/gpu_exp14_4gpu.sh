#!/bin/bash
# round 2, visit 14 (4 GPUs): host-buffer call at N = 4 -- decisions as bits (new multi-rank default) against arrays copied as they are
O=gpurun_out; mkdir -p $O
N=${1:-4}
nproc > $O/nproc_${N}gpu.txt; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA" >> $O/nproc_${N}gpu.txt
run() {  # name, port, extra env...
  local name=$1 port=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-methods > $O/bench_${N}gpu_$name.json 2> $O/bench_${N}gpu_$name.err
  echo "$name rc=$?"
  python - "$name" "$N" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/bench_%sgpu_%s.json' % (sys.argv[2], sys.argv[1])).read().strip().splitlines()[-1])
e=d['e2e']
print(sys.argv[1], 'value', round(d['value'],1), 'e2e', round(e['value'],2), 'ceiling(equal shards)', round(e['copy_ceiling']['equal_shards_ceiling_info_gbps'],1), {k: e['host_path'][k] for k in ('threads','llr_nibbles_in','decision_bits_out')})
PY
}
run default 29631 LDPC_B200_DUMMY=0
run raw 29632 LDPC_B200_HOST_THREADS=0
run default2 29633 LDPC_B200_DUMMY=0
run raw2 29634 LDPC_B200_HOST_THREADS=0
