#!/bin/bash
# Round 2: the default workload at N = 4 (weak scaling), final build
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 3 --warmup 3 --no-extra > gpurun_out/r2n4_bench_n4.json 2> gpurun_out/r2n4_bench_n4.err; cut -c1-300 gpurun_out/r2n4_bench_n4.json
