#!/bin/bash
# Round-2 GPU visit 7 (2 GPUs): library collective under torchrun, two-GPU tests, copy ceiling and e2e at N = 2, BF stage timing.
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_2gpu.txt 2>&1; lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA" >> $O/topo_2gpu.txt
( timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_threads.py -m gpu -q > $O/pytest_gpu_r02g_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02g_2gpu.log )
tail -5 $O/pytest_gpu_r02g_2gpu.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/copy_probe.py > $O/copy_probe_2gpu.json 2> $O/copy_probe_2gpu.err; cat $O/copy_probe_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_r02_2gpu.json 2> $O/bench_r02_2gpu.err; echo "bench rc=$?"; tail -3 $O/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_2gpu.json'))
for k in ('value','n_gpus','ms_per_step','gpu_launches','collective'): print(k, d[k])
print('e2e', d['e2e']['value'], d['e2e']['copy_ceiling'], d['e2e']['host_path'])
print('variants', {k:v.get('value') for k,v in d['e2e']['variants'].items()})
print('packed', d['e2e_packed_layouts'].get('value'), 'sim', d['e2e_simulate_round'].get('value'))
PY
LDPC_B200_STAGE_IN=1 LDPC_B200_STAGE_OUT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-methods > $O/bench_r02_2gpu_staged.json 2> $O/bench_r02_2gpu_staged.err; echo "bench staged rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_2gpu_staged.json'))
print('staged at N=2: e2e', d['e2e']['value'], d['e2e']['host_path'])
PY
timeout 200 python tools/nms_ab.py 3,4,2,5 1024 3.6
