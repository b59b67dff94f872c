#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ctas_summary.txt
for wl in ${WORKLOADS:-bookshelf_K65536}; do
for c in ${CTAS:-7 6 5 4 3 2}; do
  PG_ROLLOUT_CTAS=$c timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 1 --workload $wl > gpurun_out/ctas.log 2>&1
  echo "$wl ctas=$c $(tail -1 gpurun_out/ctas.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"])' 2>&1 | tail -1)" | tee -a gpurun_out/ctas_summary.txt
done; done
