#!/bin/bash
# Round 2, call c: where a rollout's cycles go (phase clocks), thread counts with rows in local memory vs on chip
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in clkT64C12 clkT256C0 clkT32C24; do
  PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so timeout 300 python tools/gpu_phase_clocks.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2c_phase.jsonl
done
PREGRASP_B200_LIB=$V/libpregrasp_b200_clkT32C24.so timeout 300 python tools/gpu_phase_clocks.py bookshelf 50 360 2>&1 | tail -1 | tee -a gpurun_out/r2c_phase.jsonl
for v in T256C0 T64C12nofma; do
  PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2c_phase.jsonl
done
