#!/bin/bash
# One gpurun call: GPU tests + rank-of-N timings. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for wr in "1 0" "2 0" "4 1" "8 3" "8 0"; do timeout 120 python tools/issue_time.py $wr 20 5; done > gpurun_out/issue_time.txt 2>&1
for c in 1 2 3 4; do timeout 120 python tools/issue_time.py 1 0 20 $c; done >> gpurun_out/issue_time.txt 2>&1
cat gpurun_out/issue_time.txt
