#!/bin/bash
# Round 2, call d: ncu of the on-chip-rows rollout kernel, one wave (K = 148 x 64), full set + source counters
mkdir -p gpurun_out
timeout 200 python tools/gpu_phys_time.py bookshelf 50 9472 1 > gpurun_out/r2d_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2d_T64C12 python tools/gpu_phys_time.py bookshelf 50 9472 1 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_plain.log; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/
