#!/bin/bash
# times "<library> <workload>" pairs given as arguments lib:workload (library relative to synthesize_pregrasp_b200/)
mkdir -p gpurun_out
for pair in "$@"; do
  so=synthesize_pregrasp_b200/${pair%%:*}; wl=${pair#*:}
  PREGRASP_B200_LIB=$PWD/$so timeout ${TMO:-600} python bench.py --steps ${STEPS:-2} --warmup 1 --cpu-sample 1 --workload $wl > gpurun_out/variant.log 2>&1
  echo "$so $wl $(tail -1 gpurun_out/variant.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"])' 2>&1 | tail -1)" | tee -a gpurun_out/variants_summary.txt
done
