#!/bin/bash
# N-GPU weak-scaling run of bench.py (torchrun, NCCL); N from $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --cpu-sample 2 > gpurun_out/scale_n$N.log 2>&1
tail -1 gpurun_out/scale_n$N.log | cut -c1-400
