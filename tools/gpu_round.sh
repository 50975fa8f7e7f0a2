#!/bin/bash
mkdir -p gpurun_out
python tools/make_gpu_golden.py > gpurun_out/make_gpu_golden.log 2>&1; tail -3 gpurun_out/make_gpu_golden.log
cp gpurun_out/gpu_frames.json tests/golden/gpu_frames.json
for c in 5 1 2 3 4; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_c${c}.json 2> gpurun_out/bench_c${c}.err
  tail -3 gpurun_out/bench_c${c}.err
done
timeout 300 python bench.py --impl reference --config 5 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
