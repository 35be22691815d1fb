#!/bin/bash
# round 2, visit 20 (4-GPU box, 32 host CPUs): what the host's memory system gives the staging passes and a streaming copy there
O=gpurun_out; mkdir -p $O
timeout 300 python tools/hostpack_bench.py 512 32 > $O/hostpack_bench_4gpu_box.log 2>&1; cat $O/hostpack_bench_4gpu_box.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 tools/copy_probe.py > $O/copy_probe_4gpu.json 2> $O/copy_probe_4gpu.err; cat $O/copy_probe_4gpu.json | head -c 1500
