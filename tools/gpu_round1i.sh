#!/bin/bash
# A/B of one solver change on the last GPU minutes: full GPU tests, then the two headline workloads (short)
mkdir -p gpurun_out
timeout 70 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for wl in bookshelf_K65536 plate_K65536; do
  timeout 45 python bench.py --steps 2 --warmup 2 --cpu-sample 1 --workload $wl > gpurun_out/i_$wl.log 2>&1
  echo "$wl $(tail -1 gpurun_out/i_$wl.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"], d["value"])' 2>&1 | tail -1)"
done
