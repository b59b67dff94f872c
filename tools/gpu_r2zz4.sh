#!/bin/bash
# last GPU seconds of round 2: default workload (without the plate extras) and plate_K360 of the final build
mkdir -p gpurun_out
timeout 40 python bench.py --steps 3 --warmup 3 --cpu-sample 8 --no-extra > gpurun_out/r2zz4_bench_bookshelf.log 2>&1; tail -1 gpurun_out/r2zz4_bench_bookshelf.log | cut -c1-300
timeout 25 python bench.py --workload plate_K360 --steps 3 --warmup 3 --cpu-sample 8 > gpurun_out/r2zz4_bench_plate_K360.log 2>&1; tail -1 gpurun_out/r2zz4_bench_plate_K360.log | cut -c1-300
