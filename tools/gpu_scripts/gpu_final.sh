#!/bin/bash
# Round-end evidence on one B200: parity tests, headline bench, ncu launch list of the same bench command.
O=gpurun_out; TAG=${1:-final}
( timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log )
tail -3 $O/pytest_gpu_$TAG.log
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; echo "bench ref rc=$?"
# launch list (cold-cache, serialised): the kernel's SHARE of the step must agree with the CUDA-event numbers above
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-methods > $O/ncu_bench_$TAG.log 2>&1; echo "ncu launch list rc=$?"
python -c "import smoke" 2>/dev/null; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
