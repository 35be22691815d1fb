#!/bin/bash
# Round-2 GPU visit 24: the rest of the GPU suite on the build with the aligned bit expansion (visit 23 ran the host-path files)
mkdir -p gpurun_out
( timeout 170 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_decode.py --deselect tests/test_gpu_host_shim.py --deselect tests/test_gpu_threads.py --deselect tests/test_gpu_dropin_reference_frontend.py > gpurun_out/pytest_gpu_exp24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_exp24.log ); tail -4 gpurun_out/pytest_gpu_exp24.log
