#!/bin/bash
# usage: bash tools/gpu_multi_all.sh N   (under gpurun --gpus N): the bench line with both exchanges + the in-process device group
N=$1
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_n${N}${tag}.json 2> gpurun_out/bench_n${N}${tag}.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}${tag}.json').read().strip().splitlines()[-1])
    print('N=$N$tag ms',round(d['ms_per_step'],4),'fps',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'stream',round(d.get('e2e_streaming',{}).get('value',0),1),'golden',d['matches_golden']['color'])
except Exception as e:
    print('N=$N$tag failed', e)
PY
}
run ""
run _nccl --composite nccl
timeout 200 python tools/group_bench.py $N 2>&1 | tail -1 > gpurun_out/group_n${N}_host.json; cut -c1-400 gpurun_out/group_n${N}_host.json
timeout 200 python tools/group_bench.py $N --peer 2>&1 | tail -1 > gpurun_out/group_n${N}_peer.json; cut -c1-330 gpurun_out/group_n${N}_peer.json
timeout 200 python tools/group_bench.py $N --nccl 2>&1 | tail -1 > gpurun_out/group_n${N}_nccl.json; cut -c1-330 gpurun_out/group_n${N}_nccl.json
