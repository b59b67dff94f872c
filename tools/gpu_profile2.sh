#!/bin/bash
mkdir -p gpurun_out
K=${1:-4096}
python bench.py --steps 1 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel" -s 1 -c 1 -o gpurun_out/prof_r1b \
    python bench.py --steps 1 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_ncu3.log 2>&1
tail -2 gpurun_out/prof_ncu3.log
