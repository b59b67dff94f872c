#!/bin/bash
# Round 2, call y: box kernel with / without FMA contraction: recorded trajectories per row + timing
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for lib in "" $V/libpregrasp_b200_nofma.so; do
  export PREGRASP_B200_LIB=$lib; [ -z "$lib" ] && unset PREGRASP_B200_LIB
  timeout 300 python tools/gpu_rec_rows.py cardboard_20.0 keyboard_20.0 foodbox_20.0 wall_laptop_20.0 keyboard_scale_20.0 2>&1 | grep "vs recording" | tee -a gpurun_out/r2y.log
  timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 0 bench 2>&1 | tail -1 | tee -a gpurun_out/r2y.log
done
