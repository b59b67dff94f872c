#!/bin/bash
# ncu evidence: launch list of one short bench run + one full capture of the physics and cost kernels
mkdir -p gpurun_out
K=${1:-8192}
python bench.py --steps 2 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_ncu1.log 2>&1
python bench.py --steps 1 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rollout_kernel|pcn_tc_kernel" -s 2 -c 2 -o gpurun_out/prof_r1c \
    python bench.py --steps 1 --warmup 3 --K $K --cpu-sample 1 > gpurun_out/prof_ncu2.log 2>&1
ls -la gpurun_out | tail; tail -3 gpurun_out/prof_ncu2.log
