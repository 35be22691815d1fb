for eb in 3.0 3.3 3.6 3.9 4.2; do echo "== Eb/N0 $eb"; python tools/nms_ab.py 1,2,4 1024 $eb; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_pair -c 1 -f -o gpurun_out/prof_oms_v6 python tools/prof_decode.py 1 1024 > gpurun_out/ncu_oms_v6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize -c 1 -f -o gpurun_out/prof_fin4_v6 python tools/prof_decode.py 4 1024 > gpurun_out/ncu_fin4_v6.log 2>&1
