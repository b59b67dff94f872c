#!/bin/bash
# Round 2, call i: double-buffered cost kernel (parity + time), new tests (update selection, device vs host, recorded rows)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -s -k "cost or score or csg" 2>&1 | tail -8 | tee gpurun_out/r2i_pytest_cost.log
timeout 300 python bench.py --steps 2 --warmup 2 --no-extra --cpu-sample 16 > gpurun_out/r2i_bench.log 2>&1; tail -1 gpurun_out/r2i_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['roofline_cost']['frac'])"
timeout 1200 python -m pytest tests -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/r2i_pytest_all.log
