#!/bin/bash
# Round 2, first GPU call: micro-benchmarks that size the new rollout kernel + re-baseline of the round-1 final build.
#   gpurun --timeout 900 -- 'bash tools/gpu_r2a.sh'
mkdir -p gpurun_out
timeout 120 tools/ubench/ubench > gpurun_out/r2a_ubench.json 2> gpurun_out/r2a_ubench.err; cat gpurun_out/r2a_ubench.json; tail -3 gpurun_out/r2a_ubench.err
timeout 400 python -m pytest tests -q -m gpu --durations=6 2>&1 | tail -12 | tee gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_default.log 2>&1; tail -1 gpurun_out/r2a_bench_default.log | cut -c1-400
for wl in plate_K65536 plate_K360; do
  timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 64 --workload $wl > gpurun_out/r2a_bench_$wl.log 2>&1
  tail -1 gpurun_out/r2a_bench_$wl.log | cut -c1-400
done
