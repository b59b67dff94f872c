#!/bin/bash
# Round 2, call g: periodic settle re-deal (fixed), launch-sized solver slices (L1 kept for small batches), hull phase clocks
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in RG100 RG50 RTS; do
  PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so timeout 200 python tools/gpu_phys_time.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
done
for v in default RTS; do
  if [ $v = default ]; then unset PREGRASP_B200_LIB; else export PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so; fi
  timeout 200 python tools/gpu_phys_time.py plate 20 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
  timeout 200 python tools/gpu_phys_time.py bookshelf 50 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
  timeout 200 python tools/gpu_phys_time.py plate 20 65536 2 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
done
PREGRASP_B200_LIB=$V/libpregrasp_b200_clkhull.so timeout 300 python tools/gpu_phase_clocks.py plate 20 65536 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
PREGRASP_B200_LIB=$V/libpregrasp_b200_clkhull.so timeout 300 python tools/gpu_phase_clocks.py plate 20 360 2>&1 | tail -1 | tee -a gpurun_out/r2g_variants.jsonl
