#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 "$@" > gpurun_out/bench_n2_$tag.json 2> gpurun_out/bench_n2_$tag.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n2_$tag.json').read().strip().splitlines()[-1])
print('$tag','ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],1))
"; }
run p2p
run nccl --composite nccl
run imm --immediate
