#!/bin/bash
# Round-2 GPU visit 8: BF stage (class-wise compaction, conditional flips), PRMT epilogue, staged e2e vs chunk size / NT stores.
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity_at_scale.py tests/test_gpu_random_configs.py -m gpu -q > $O/pytest_gpu_r02h.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02h.log )
tail -6 $O/pytest_gpu_r02h.log
L=$O/nms_ab_exp8.log; : > $L
timeout 300 python tools/nms_ab.py 0,1,2,3,4,5 1024 3.6 >> $L 2>&1
timeout 300 python tools/nms_ab.py 0 2048 3.6 >> $L 2>&1
cat $L
timeout 300 python tools/e2e_chunks.py 2048 > $O/e2e_chunks_exp8.log 2>&1
LDPC_B200_PACK_NT=0 timeout 300 python tools/e2e_chunks.py 2048 >> $O/e2e_chunks_exp8.log 2>&1
cat $O/e2e_chunks_exp8.log
