#!/bin/bash
# Round 2, call b: CTA size x on-chip row capacity sweep of the box rollout kernel (bookshelf K = 65536), speed + digest.
mkdir -p gpurun_out
V=synthesize_pregrasp_b200/variants
for lib in default $V/libpregrasp_b200_T96C7.so $V/libpregrasp_b200_T128C4.so $V/libpregrasp_b200_T32C24.so $V/libpregrasp_b200_T128C0.so $V/libpregrasp_b200_T448C0.so; do
  if [ $lib = default ]; then unset PREGRASP_B200_LIB; else export PREGRASP_B200_LIB=$PWD/$lib; fi
  timeout 200 python tools/gpu_phys_time.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2b_variants.jsonl
done
unset PREGRASP_B200_LIB
timeout 100 python tools/gpu_phys_time.py bookshelf 50 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2b_variants.jsonl
timeout 100 python tools/gpu_phys_time.py plate 20 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2b_variants.jsonl
timeout 200 python tools/gpu_phys_time.py plate 20 65536 2>&1 | tail -1 | tee -a gpurun_out/r2b_variants.jsonl
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2b_pytest.log
