#!/bin/bash
# usage: bash tools/gpu_multi_host.sh N [--peer] -- bench at N ranks (e2e composed in host memory) + the in-process device group, three presents
N=$1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/bench_n$N.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1),d['e2e'].get('composite'),d['e2e']['image_matches_resident_path'],'stream',round(d['e2e_streaming']['value'],1),d['e2e_streaming']['image_matches_resident_path'],d['matches_golden']['color'])
PY
for mode in "" $2; do
  timeout 200 python tools/group_bench.py $N $mode 2>&1 | tail -1 > gpurun_out/group_n${N}${mode/--/_}.json; cut -c1-300 gpurun_out/group_n${N}${mode/--/_}.json; echo
done
