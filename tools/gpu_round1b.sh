#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 1 --K 8192 --cpu-sample 2 > gpurun_out/bench_K8192.log 2>&1
timeout 1500 python bench.py --steps 3 --warmup 3 --cpu-sample 4 > gpurun_out/bench_default.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_K8192.log; tail -2 gpurun_out/bench_default.log
