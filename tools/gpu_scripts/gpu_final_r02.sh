#!/bin/bash
# Round-2 end evidence on one B200: parity tests, headline bench + reference arm, ncu launch list of the same bench command,
# ncu --set full of the NMS kernel (roofline inputs, stalls) and of every kernel, smoke().
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02_final.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02_final.log )
tail -5 $O/pytest_gpu_r02_final.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r02_ref_final.json 2> $O/bench_r02_final.err; echo "bench ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_r02_final.json 2>> $O/bench_r02_final.err; echo "bench rc=$?"; tail -2 $O/bench_r02_final.err
# launch list (cold-cache, serialised): the kernel's SHARE of the step must agree with the CUDA-event numbers above
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-methods --groups 2048 > $O/ncu_bench_r02_final.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_pair -c 1 -f -o $O/nms_r02_final python tools/nms_ab.py 0 1024 3.6 > $O/ncu_nms_r02_final.log 2>&1; echo "ncu nms rc=$?"
timeout 900 ncu --set full --clock-control none -c 60 -f -o /tmp/all_r02_final python tools/prof_all.py > $O/ncu_all_r02_final.log 2>&1; echo "ncu all rc=$?"
# the all-kernels report is too big to bring back (gpurun_out is capped at 64 MiB): its raw page goes home as CSV
ncu -i /tmp/all_r02_final.ncu-rep --page raw --csv > $O/all_r02_final_raw.csv 2>/dev/null; rm -f /tmp/all_r02_final.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/nms_ab.py 0,1,2,3,4,5 1024 3.6 > $O/nms_ab_r02_final.log 2>&1; cat $O/nms_ab_r02_final.log
