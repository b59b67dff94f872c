#!/bin/bash
# last GPU seconds of round 2: plate bench line of the final build (hull kernel with the double-precision GJK tolerances)
mkdir -p gpurun_out
timeout 95 python bench.py --workload plate_K65536 --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/r2zz3_bench_plate.log 2>&1; tail -1 gpurun_out/r2zz3_bench_plate.log | cut -c1-400
