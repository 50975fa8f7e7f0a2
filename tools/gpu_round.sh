#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python tools/issue_time.py 1 0 20 5 | sed 's/world 1 rank 0 stripe 270//'; env "$@" timeout 120 python tools/issue_time.py 8 3 20 5 | sed 's/world 8 rank 3 stripe 34//'; }
{
run MLV_FRONT_STREAMS=1
run MLV_FRONT_STREAMS=1 MLV_EXP_SKIP_TAIL=1
run MLV_FRONT_STREAMS=2 MLV_EXP_SKIP_TAIL=1
run MLV_FRONT_STREAMS=8 MLV_EXP_SKIP_TAIL=1
} > gpurun_out/knobs.txt 2>&1
cat gpurun_out/knobs.txt
