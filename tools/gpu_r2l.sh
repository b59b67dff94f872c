#!/bin/bash
# Round 2, call l: producer/consumer cost kernel: parity + time
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "cost or score or csg or end_to_end or update_selection" 2>&1 | tail -5 | tee gpurun_out/r2l_pytest_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 8192 3 2>&1 | tail -1 | tee gpurun_out/r2l_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 131072 2 2>&1 | tail -1 | tee -a gpurun_out/r2l_cost.log
