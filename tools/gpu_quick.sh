#!/bin/bash
# GPU tests + the two headline workloads, short
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for wl in bookshelf_K65536 plate_K65536; do
  timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 1 --workload $wl > gpurun_out/quick_$wl.log 2>&1
  echo "$wl $(tail -1 gpurun_out/quick_$wl.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"], d["value"])' 2>&1 | tail -1)"
done
