#!/bin/bash
# first GPU pass: tests, smoke, a small and the default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --K 8192 --cpu-sample 2 > gpurun_out/bench_K8192.log 2>&1
timeout 1500 python bench.py --steps 3 --warmup 3 --cpu-sample 4 > gpurun_out/bench_default.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -2 gpurun_out/bench_K8192.log; tail -2 gpurun_out/bench_default.log
