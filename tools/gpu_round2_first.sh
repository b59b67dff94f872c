#!/bin/bash
# First GPU call of the next round: the last two solver changes of round 1 (per-island solves, pair-cache order with the first
# wall ahead of the fingertip pairs) went in after the GPU minutes ran out -- re-measure everything on that build.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -12 | tee gpurun_out/r2_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2_bench_default.log 2>&1; tail -1 gpurun_out/r2_bench_default.log | cut -c1-250
for wl in plate_K65536 plate_K360 plate_Q30xK360 waterbottle_K65536; do
  timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 64 --workload $wl > gpurun_out/r2_bench_$wl.log 2>&1
  tail -1 gpurun_out/r2_bench_$wl.log | cut -c1-250
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.log 2>&1; tail -1 gpurun_out/r2_bench_reference.log | cut -c1-250
PR_K=160 timeout 600 python tools/gpu_parity_report.py > gpurun_out/r2_parity_kernel_vs_oracle_K160.json 2>gpurun_out/r2_parity.err; tail -30 gpurun_out/r2_parity_kernel_vs_oracle_K160.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bookshelf_K65536.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/r2_ncu.log 2>&1; wc -l gpurun_out/r2_launches_bookshelf_K65536.csv
for WL in bookshelf_K65536 plate_K65536; do
  timeout 400 ncu --clock-control none -k regex:rollout_kernel -c 1 \
    --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section SchedulerStats --section Occupancy --section SpeedOfLight --section LaunchStats \
    --csv --page raw --log-file gpurun_out/r2_sections_${WL}.csv python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL > gpurun_out/r2_ncu_$WL.log 2>&1
done
ls -la gpurun_out/
