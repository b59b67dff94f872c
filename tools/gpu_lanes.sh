#!/bin/bash
# small-batch mapping: rollouts per warp (PG_LANES) for the reference-sized workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for wl in plate_K360 plate_Q30xK360; do
  for ln in 0 32 ${EXTRA_LANES}; do
    PG_LANES=$ln timeout 300 python bench.py --steps 3 --warmup 2 --cpu-sample 1 --workload $wl > gpurun_out/variant.log 2>&1
    echo "lanes=$ln $wl $(tail -1 gpurun_out/variant.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["stage_ms"])' 2>&1 | tail -1)" | tee -a gpurun_out/lanes_summary.txt
  done
done
