#!/bin/bash
# memory / stall picture of the rollout kernel at the FULL bench size (K = 65536): one launch, selected sections
mkdir -p gpurun_out
WL=${WL:-bookshelf_K65536}
timeout 600 python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL > gpurun_out/p3_plain.log 2>&1 || exit 1
tail -1 gpurun_out/p3_plain.log | cut -c1-300
timeout 1500 ncu --clock-control none -k regex:rollout_kernel -c 1 \
  --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section SchedulerStats --section Occupancy --section SpeedOfLight --section LaunchStats \
  --csv --page raw --log-file gpurun_out/p3_${WL}_raw.csv python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL > gpurun_out/p3_ncu.log 2>&1
tail -3 gpurun_out/p3_ncu.log | cut -c1-200
ls -la gpurun_out/p3_${WL}_raw.csv
