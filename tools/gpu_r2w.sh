#!/bin/bash
# Round 2, call w: the build with the idle fingertips' manifolds in the dispatcher order: GPU tests + physics timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; tail -15 gpurun_out/r2w_pytest.log
timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 0 bench 2>&1 | tail -1 | tee -a gpurun_out/r2w.jsonl
timeout 300 python tools/gpu_phys_time.py plate 20 65536 2 0 bench 2>&1 | tail -1 | tee -a gpurun_out/r2w.jsonl
