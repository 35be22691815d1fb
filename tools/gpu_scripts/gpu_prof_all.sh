#!/bin/bash
# ncu --set full of every kernel; the (large) report stays on the box, only the raw CSV page comes back
O=gpurun_out; TAG=${1:-r01}
timeout 300 python tools/prof_all.py 296 > $O/prof_all_plain_$TAG.log 2>&1; echo "plain rc=$?"; tail -3 $O/prof_all_plain_$TAG.log
timeout 1500 ncu --set full --clock-control none --import-source off -k regex:'decode_pair|finalize|generate|encode|count_errors|info_bits|quantize|group_hist' -f -o /tmp/prof_all_$TAG python tools/prof_all.py 296 > $O/ncu_all_$TAG.log 2>&1; echo "ncu rc=$?"; tail -2 $O/ncu_all_$TAG.log
ncu -i /tmp/prof_all_$TAG.ncu-rep --page raw --csv > $O/prof_all_$TAG.csv 2>/dev/null
ls -la /tmp/prof_all_$TAG.ncu-rep $O/prof_all_$TAG.csv
