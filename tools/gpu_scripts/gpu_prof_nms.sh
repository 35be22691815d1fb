#!/bin/bash
O=gpurun_out; TAG=${1:-nms}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pair -c 1 -f -o /tmp/prof_nms_$TAG python tools/prof_decode.py 0 1024 > $O/ncu_nms_$TAG.log 2>&1
ncu -i /tmp/prof_nms_$TAG.ncu-rep --page raw --csv > $O/prof_nms_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_nms_$TAG.ncu-rep --page source --csv > $O/prof_nms_${TAG}_source.csv 2>/dev/null
ls -la /tmp/prof_nms_$TAG.ncu-rep $O/prof_nms_${TAG}_*.csv
