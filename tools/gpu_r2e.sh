#!/bin/bash
# Round 2, call e: lean (register-resident) sweep -- CTA size x on-chip capacity sweep, digest check, phase clocks
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in L_T64C12 L_T128C5 L_T256C0 L_T96C7 L_T192C2; do
  PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2e_variants.jsonl
done
PREGRASP_B200_LIB=$V/libpregrasp_b200_L_clkT64C12.so timeout 300 python tools/gpu_phase_clocks.py bookshelf 50 9472 2>&1 | tail -1 | tee -a gpurun_out/r2e_variants.jsonl
for v in L_T64C12 L_T256C0; do
  PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so timeout 300 python tools/gpu_phys_time.py bookshelf 50 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2e_variants.jsonl
done
