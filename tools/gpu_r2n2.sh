#!/bin/bash
# Round 2: two GPUs of one box: the two-devices-in-one-process test, and the default workload at N = 2 (weak scaling)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "two_devices" -rA > gpurun_out/r2n2_pytest.log 2>&1; tail -4 gpurun_out/r2n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > gpurun_out/r2n2_bench_n2.json 2> gpurun_out/r2n2_bench_n2.err; cut -c1-300 gpurun_out/r2n2_bench_n2.json
