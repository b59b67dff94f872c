#!/bin/bash
# GPU tests + plate (hull) workload bench + default workload bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 2 --warmup 1 --workload plate_K65536 --K 8192 --cpu-sample 2 > gpurun_out/bench_plate_K8192.log 2>&1
tail -2 gpurun_out/bench_plate_K8192.log
timeout 1500 python bench.py --steps 3 --warmup 3 --workload plate_K65536 --cpu-sample 4 > gpurun_out/bench_plate.log 2>&1
tail -2 gpurun_out/bench_plate.log
timeout 1500 python bench.py --steps 3 --warmup 3 --cpu-sample 4 > gpurun_out/bench_default.log 2>&1
tail -2 gpurun_out/bench_default.log
