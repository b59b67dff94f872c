#!/bin/bash
# source-level ncu capture of the hull (plate) rollout kernel at K = 8192
mkdir -p gpurun_out
WL=${WL:-plate_K65536}; K=${K:-8192}
timeout 600 python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL --K $K > gpurun_out/p4_plain.log 2>&1 || exit 1
tail -1 gpurun_out/p4_plain.log | cut -c1-200
timeout 2000 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/p4_${WL} -f \
  python bench.py --steps 1 --warmup 1 --cpu-sample 1 --workload $WL --K $K > gpurun_out/p4_ncu.log 2>&1
tail -2 gpurun_out/p4_ncu.log | cut -c1-200; ls -la gpurun_out/p4_*
