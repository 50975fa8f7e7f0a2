#!/bin/bash
# One GPU call that produces every raw capture tools/make_profiles.py turns into profiles/r02_*:
#   ncu launch list of one frame (cold-cache, serialised), ncu --set full of the visible draw + the first hidden draw,
#   device timelines of a recorded frame (1 GPU, one rank of 8), bench lines of configs 1-5, the reference arm, sanitizers.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size
# a frame of config 5 in immediate mode = 58 launches (clear, 8 x {k_vertex k_front k_front_clip k_back k_bin_scan k_fill k_tile}, resolve)
timeout 600 ncu --metrics $M --clock-control none -s 116 -c 58 --csv --log-file gpurun_out/launches.csv python tools/profile_rank.py 1 0 3 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 116 -c 15 -f -o gpurun_out/full python tools/profile_rank.py 1 0 3 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -s 150 -c 50 --csv --log-file gpurun_out/launches_w8_r3.csv python tools/profile_rank.py 8 3 4 > gpurun_out/ncu_w8.log 2>&1
timeout 120 python tools/timeline_frame.py 1 0 5 gpurun_out/timeline_config5_n1.json > gpurun_out/timeline_config5_n1.txt 2>&1
timeout 120 python tools/timeline_frame.py 8 3 5 gpurun_out/timeline_config5_rank3of8.json > gpurun_out/timeline_config5_rank3of8.txt 2>&1
timeout 120 python tools/timeline_frame.py 1 0 2 gpurun_out/timeline_config2_n1.json > gpurun_out/timeline_config2_n1.txt 2>&1
timeout 120 python tools/launch_table.py 1 0 5 > gpurun_out/launch_table_config5_n1.txt 2>&1
[ -n "$QUICK" ] || timeout 120 python tools/trace_frame.py 5 gpurun_out/trace_config5.json > /dev/null 2>&1
for c in 5 1 2 3 4; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_c${c}.json 2> gpurun_out/bench_c${c}.err
  tail -2 gpurun_out/bench_c${c}.err
done
timeout 300 python bench.py --impl reference --steps ${REF_STEPS:-20} --warmup ${REF_WARMUP:-3} > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 compute-sanitizer --tool memcheck python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1; tail -2 gpurun_out/sanitizer_memcheck.log
[ -n "$QUICK" ] || timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck.log
[ -n "$QUICK" ] || timeout 300 compute-sanitizer --tool synccheck python __graft_entry__.py smoke > gpurun_out/sanitizer_synccheck.log 2>&1; tail -2 gpurun_out/sanitizer_synccheck.log
ls -la gpurun_out/full.ncu-rep gpurun_out/launches.csv
