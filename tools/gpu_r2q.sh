#!/bin/bash
# Round 2, call q: ncu evidence of the build: launch list of the default bench command, full-set captures of rollout_kernel<box>
# (bench size) and pcn_tc_kernel, each right after the same command exited 0 without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extra --cpu-sample 1"
$CMD > gpurun_out/r2q_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_bookshelf_K65536.csv $CMD > gpurun_out/r2q_ncu1.log 2>&1
wc -l gpurun_out/r2q_launches_bookshelf_K65536.csv
python tools/gpu_phys_time.py bookshelf 50 65536 1 > gpurun_out/r2q_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2q_rollout_box python tools/gpu_phys_time.py bookshelf 50 65536 1 > gpurun_out/r2q_ncu2.log 2>&1
tail -2 gpurun_out/r2q_ncu2.log
python tools/gpu_cost_time.py bookshelf 16384 1 > gpurun_out/r2q_plain3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcn_tc_kernel -s 1 -c 1 -o gpurun_out/r2q_cost python tools/gpu_cost_time.py bookshelf 16384 1 > gpurun_out/r2q_ncu3.log 2>&1
tail -2 gpurun_out/r2q_ncu3.log; ls -la gpurun_out | tail -8
