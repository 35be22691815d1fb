#!/bin/bash
# Round-2, closing check: the whole GPU suite and smoke() on the final tree
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_r02_final3.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r02_final3.log )
tail -5 $O/pytest_gpu_r02_final3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
