#!/bin/bash
# round 2, visit 13 (2 GPUs): host-buffer call at N = 2 -- library default (arrays copied as they are) against decisions as bits
# and against the hybrid path with each rank's share of the host threads
O=gpurun_out; mkdir -p $O
run() {  # name, port, extra env...
  local name=$1 port=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-methods > $O/bench_2gpu_$name.json 2> $O/bench_2gpu_$name.err
  echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/bench_2gpu_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
e=d['e2e']
print(sys.argv[1], 'value', round(d['value'],1), 'e2e', round(e['value'],2), 'ceiling(equal shards)', round(e['copy_ceiling']['equal_shards_ceiling_info_gbps'],1), e['host_path'])
PY
}
run default 29621 LDPC_B200_DUMMY=0
run bits_out 29622 LDPC_B200_STAGE_IN=0 LDPC_B200_STAGE_OUT=1
run hybrid 29623 LDPC_B200_STAGE_IN=1 LDPC_B200_STAGE_OUT=1
run default2 29624 LDPC_B200_DUMMY=0
run bits_out2 29625 LDPC_B200_STAGE_IN=0 LDPC_B200_STAGE_OUT=1
run hybrid2 29626 LDPC_B200_STAGE_IN=1 LDPC_B200_STAGE_OUT=1
