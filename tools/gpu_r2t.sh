#!/bin/bash
# Round 2, call t: cost kernel with the A operands in tensor memory: parity + time
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "cost or score or csg or end_to_end or update_selection or step_single" 2>&1 | tail -5 | tee gpurun_out/r2t_pytest_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 16384 3 2>&1 | tail -1 | tee gpurun_out/r2t_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 131072 2 2>&1 | tail -1 | tee -a gpurun_out/r2t_cost.log
timeout 200 python tools/gpu_cost_time.py waterbottle 16384 2 2>&1 | tail -1 | tee -a gpurun_out/r2t_cost.log
