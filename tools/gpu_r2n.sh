#!/bin/bash
# Round 2, call n: timing ablations of the cost kernel (results of the ablated builds are wrong by construction)
mkdir -p gpurun_out
V=$PWD/synthesize_pregrasp_b200/variants
for v in default NG2 noMMA4 noMMA3 noMMA34 noEPI3 noEPI3noMMA4 noGEOM; do
  if [ $v = default ]; then unset PREGRASP_B200_LIB; else export PREGRASP_B200_LIB=$V/libpregrasp_b200_$v.so; fi
  echo -n "$v " | tee -a gpurun_out/r2n_ablate.log
  timeout 100 python tools/gpu_cost_time.py bookshelf 16384 2 2>&1 | tail -1 | tee -a gpurun_out/r2n_ablate.log
done
