#!/bin/bash
# bench.py stage timing as a median of three samples: one short default-workload run
mkdir -p gpurun_out
timeout 30 python bench.py --steps 2 --warmup 3 --cpu-sample 4 --no-extra > gpurun_out/r2zz5_bench_bookshelf.log 2>&1; tail -1 gpurun_out/r2zz5_bench_bookshelf.log | cut -c1-200; grep -o '"stage_ms": {[^}]*}' gpurun_out/r2zz5_bench_bookshelf.log
