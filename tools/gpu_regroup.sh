#!/bin/bash
# settle-phase regrouping: correctness test, then timings with it on / off (PG_REGROUP=0) and with CTA barriers in the box kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "regrouping or random_batch or multi_problem" 2>&1 | tail -5
for wl in bookshelf_K65536 plate_K65536; do
  for rg in 1 0; do
    PG_REGROUP=$rg timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 1 --workload $wl > gpurun_out/variant.log 2>&1
    echo "regroup=$rg $wl $(tail -1 gpurun_out/variant.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"])' 2>&1 | tail -1)" | tee -a gpurun_out/regroup_summary.txt
  done
done
PREGRASP_B200_LIB=$PWD/synthesize_pregrasp_b200/_variants/lib_box_bar.so timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 1 --workload bookshelf_K65536 > gpurun_out/variant.log 2>&1
echo "box barriers + regroup bookshelf_K65536 $(tail -1 gpurun_out/variant.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["stage_ms"])' 2>&1 | tail -1)" | tee -a gpurun_out/regroup_summary.txt
