#!/bin/bash
# Round 2, call m: 512-producer cost kernel + closed-form prism distance field: parity + time
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x -k "cost or score or csg or end_to_end or update_selection or step_single" 2>&1 | tail -5 | tee gpurun_out/r2m_pytest_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 8192 3 2>&1 | tail -1 | tee gpurun_out/r2m_cost.log
timeout 200 python tools/gpu_cost_time.py bookshelf 131072 2 2>&1 | tail -1 | tee -a gpurun_out/r2m_cost.log
timeout 200 python tools/gpu_cost_time.py waterbottle 8192 2 2>&1 | tail -1 | tee -a gpurun_out/r2m_cost.log
timeout 200 python tools/gpu_cost_time.py plate 8192 2 2>&1 | tail -1 | tee -a gpurun_out/r2m_cost.log
