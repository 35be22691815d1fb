#!/bin/bash
# round 2, visit 11: semi-direct hybrid host path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi_exp11.txt
nproc >> gpurun_out/smi_exp11.txt
timeout 900 python tools/e2e_semidirect.py 2048 > gpurun_out/e2e_semidirect.log 2>&1; echo "semidirect rc=$?"
tail -45 gpurun_out/e2e_semidirect.log
