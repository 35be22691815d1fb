#!/bin/bash
# Round-2 GPU visit 23: the host-buffer call with pageable caller arrays (malloc / numpy, 16-byte aligned, as a reference-style caller has
# them), with and without LDPC_B200_HOST_REGISTER; then the GPU tests that go through the host path
mkdir -p gpurun_out
timeout 500 python tools/e2e_pageable.py 2048 > gpurun_out/e2e_pageable.log 2>&1; cat gpurun_out/e2e_pageable.log
( timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_host_shim.py tests/test_gpu_threads.py tests/test_gpu_dropin_reference_frontend.py -m gpu -q -x > gpurun_out/pytest_gpu_exp23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_exp23.log ); tail -4 gpurun_out/pytest_gpu_exp23.log
