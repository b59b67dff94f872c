#!/bin/bash
mkdir -p gpurun_out
for m in 2; do
  timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 $m 2>&1 | tail -1 | tee -a gpurun_out/r2s_phased.jsonl
  timeout 300 python tools/gpu_phys_time.py plate 20 65536 2 $m 2>&1 | tail -1 | tee -a gpurun_out/r2s_phased.jsonl
  timeout 300 python tools/gpu_phys_time.py bookshelf 50 8192 2 $m 2>&1 | tail -1 | tee -a gpurun_out/r2s_phased.jsonl
done
