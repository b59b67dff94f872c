#!/bin/bash
# round-end evidence on the final build with the few GPU minutes left: default bench (validates bench.py), the small-batch,
# multi-problem and waterbottle workloads, then the section capture of the hull kernel at the full bench size
mkdir -p gpurun_out
timeout 240 python bench.py > gpurun_out/h_bench_default.log 2>&1; tail -1 gpurun_out/h_bench_default.log | cut -c1-300
for wl in plate_K360 plate_Q30xK360; do
  timeout 120 python bench.py --steps 3 --warmup 3 --cpu-sample 64 --workload $wl > gpurun_out/h_bench_$wl.log 2>&1
  tail -1 gpurun_out/h_bench_$wl.log | cut -c1-300
done
timeout 200 python bench.py --steps 2 --warmup 3 --cpu-sample 16 --workload waterbottle_K65536 > gpurun_out/h_bench_waterbottle.log 2>&1
tail -1 gpurun_out/h_bench_waterbottle.log | cut -c1-300
WL=plate_K65536
timeout 300 ncu --clock-control none -k regex:rollout_kernel -c 1 \
  --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section SchedulerStats --section Occupancy --section SpeedOfLight --section LaunchStats \
  --csv --page raw --log-file gpurun_out/h_sections_${WL}.csv python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL > gpurun_out/h_ncu.log 2>&1
tail -2 gpurun_out/h_ncu.log | cut -c1-200
ls -la gpurun_out/
