#!/bin/bash
# Round 2, call r: lane-homogeneity experiment; phased (work-sorted) rollouts vs the persistent kernel: digest + time
mkdir -p gpurun_out
timeout 300 python tools/gpu_homog_test.py bookshelf 50 65536 2>&1 | tail -1 | tee gpurun_out/r2r_homog.json
for m in 1 2; do
  timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 $m 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
  timeout 300 python tools/gpu_phys_time.py plate 20 65536 2 $m 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
done
timeout 300 python tools/gpu_phys_time.py bookshelf 50 8192 2 2 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
timeout 300 python tools/gpu_phys_time.py bookshelf 50 8192 2 1 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
timeout 300 python tools/gpu_phys_time.py bookshelf 50 1000 2 2 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
timeout 300 python tools/gpu_phys_time.py bookshelf 50 1000 2 1 2>&1 | tail -1 | tee -a gpurun_out/r2r_phased.jsonl
