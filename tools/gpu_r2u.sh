#!/bin/bash
# Round 2, call u: sweep without register copies (ping-pong pipelining) + base-pointer addressing: digest + time
mkdir -p gpurun_out
timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 2>&1 | tail -1 | tee -a gpurun_out/r2u.jsonl
timeout 300 python tools/gpu_phys_time.py plate 20 65536 2 2>&1 | tail -1 | tee -a gpurun_out/r2u.jsonl
timeout 300 python tools/gpu_phys_time.py plate 20 360 3 2>&1 | tail -1 | tee -a gpurun_out/r2u.jsonl
timeout 300 python tools/gpu_phys_time.py waterbottle 20 16384 1 2>&1 | tail -1 | tee -a gpurun_out/r2u.jsonl
