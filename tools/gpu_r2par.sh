#!/bin/bash
# Round 2: kernel-vs-oracle agreement report of the final build (recorded controls + 96 random rollouts per env)
mkdir -p gpurun_out
PR_K=96 timeout 900 python tools/gpu_parity_report.py > gpurun_out/r2_parity_kernel_vs_oracle_K96.json 2> gpurun_out/r2_parity.err; tail -c 1500 gpurun_out/r2_parity_kernel_vs_oracle_K96.json
