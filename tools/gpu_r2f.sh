#!/bin/bash
# Round 2, call f: round-1 kernel design re-baselined; periodic settle regroup; no substep barriers; phase clocks of the baseline
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in default RG100 RG50 RG25 nobar; do
  if [ $v = default ]; then unset PREGRASP_B200_LIB; else export PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so; fi
  timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2f_variants.jsonl
done
PREGRASP_B200_LIB=$V/libpregrasp_b200_clkbase.so timeout 300 python tools/gpu_phase_clocks.py bookshelf 50 65536 2>&1 | tail -1 | tee -a gpurun_out/r2f_variants.jsonl
unset PREGRASP_B200_LIB
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2f_pytest.log
