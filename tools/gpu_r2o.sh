#!/bin/bash
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in A1 A2 A3 A4 A5; do
  export PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so
  echo -n "$v " | tee -a gpurun_out/r2o_ablate.log
  timeout 100 python tools/gpu_cost_time.py bookshelf 16384 2 2>&1 | tail -1 | tee -a gpurun_out/r2o_ablate.log
done
