#!/bin/bash
# Round 2, call h: all-tensor-core cost kernel (parity + time), periodic regroup as default (box + hull)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 2 --warmup 2 --no-extra --cpu-sample 16 > gpurun_out/r2h_bench.log 2>&1; tail -1 gpurun_out/r2h_bench.log | cut -c1-700
timeout 200 python tools/gpu_phys_time.py plate 20 65536 2 2>&1 | tail -1 | tee -a gpurun_out/r2h_variants.jsonl
timeout 200 python tools/gpu_phys_time.py bookshelf 50 65536 2 2>&1 | tail -1 | tee -a gpurun_out/r2h_variants.jsonl
