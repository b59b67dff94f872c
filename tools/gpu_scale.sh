#!/bin/bash
# Multi-GPU lines of round 2 (one box, N = $1 GPUs):  gpurun --gpus N -- 'bash tools/gpu_scale.sh N'
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; timeout 900 $TR bench.py --gpus $N "$@" > gpurun_out/r2_scale_n${N}_$name.json 2> gpurun_out/r2_scale_n${N}_$name.err; tail -c 700 gpurun_out/r2_scale_n${N}_$name.json; echo; tail -2 gpurun_out/r2_scale_n${N}_$name.err; }
run bookshelf_K65536 --steps 3 --warmup 3 --no-extra --cpu-sample 16
run waterbottle_K1048576 --workload waterbottle_K1048576 --steps 1 --warmup 1 --cpu-sample 8
if [ "$N" = 8 ]; then
  run dyn_sweep_P1024 --workload dyn_sweep_P1024 --steps 2 --warmup 1
  run multi_env --workload multi_env --steps 2 --warmup 1
fi
