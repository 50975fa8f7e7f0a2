#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
for wr in "1 0" "8 3" "8 0" "2 0" "4 1"; do timeout 120 python tools/issue_time.py $wr 20 5 | grep "command list"; done > gpurun_out/issue_time.txt 2>&1
for c in 1 2 3 4; do timeout 120 python tools/issue_time.py 1 0 20 $c| grep "command list"; done >> gpurun_out/issue_time.txt 2>&1
cat gpurun_out/issue_time.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size --clock-control none --csv --log-file gpurun_out/launches_w1_r0.csv python tools/profile_rank.py 1 0 2 5 > gpurun_out/ncu_w1.log 2>&1
