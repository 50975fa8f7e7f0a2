#!/bin/bash
# one GPU call: parity tests, config-5 bench line, a rank of 8 alone
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -2 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
print('c5 ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),'stream',round(d['e2e_streaming']['value'],1),d['roofline']['kernel'],round(d['roofline']['frac'],3),d['matches_golden']['color'],d['matches_golden']['depth'])
print(d['stage_ms_per_step'])
PY
(timeout 90 python tools/issue_time.py 1 0 20; timeout 90 python tools/issue_time.py 8 0 20; timeout 90 python tools/issue_time.py 8 3 20) 2>&1 | grep "command list" | tee gpurun_out/issue_time.txt
