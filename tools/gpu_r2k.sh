#!/bin/bash
# Round 2, call k: cost kernel with lean MMA issue; full GPU suite with printed agreement figures
mkdir -p gpurun_out
timeout 200 python tools/gpu_cost_time.py bookshelf 8192 3 2>&1 | tail -1 | tee gpurun_out/r2k_cost.log
timeout 200 python tools/gpu_cost_time.py waterbottle 8192 2 2>&1 | tail -1 | tee -a gpurun_out/r2k_cost.log
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" > gpurun_out/r2k_pytest_all.log; grep -E "within|arg-max|device vs host|kernel vs PyBullet|passed|failed|FAILED|Error" gpurun_out/r2k_pytest_all.log | cut -c1-330
