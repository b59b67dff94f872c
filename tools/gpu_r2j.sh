#!/bin/bash
# Round 2, call j: ncu of the tensor-core cost kernel (one launch, 8192 evaluations), full set + source
mkdir -p gpurun_out
timeout 200 python tools/gpu_cost_time.py bookshelf 8192 2 > gpurun_out/r2j_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcn_tc_kernel -s 1 -c 1 -o gpurun_out/r2j_cost python tools/gpu_cost_time.py bookshelf 8192 1 > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_plain.log; tail -2 gpurun_out/r2j_ncu.log
