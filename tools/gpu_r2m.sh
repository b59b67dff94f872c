#!/bin/bash
# Round 2: MPPI final fold parallelised: its tests, the stand-alone line, and the traffic record of the (unchanged) rollout kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mppi or update_selection or end_to_end or optimizer_api" > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
timeout 300 python bench.py --workload mppi_update_K1048576 > gpurun_out/r2m_bench_mppi_update.json 2> gpurun_out/r2m_bench_mppi_update.err; cut -c1-200 gpurun_out/r2m_bench_mppi_update.json
python tools/gpu_phys_time.py bookshelf 50 65536 1 0 bench > gpurun_out/r2m_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/r2m_rollout_box python tools/gpu_phys_time.py bookshelf 50 65536 1 0 bench > gpurun_out/r2m_ncu2.log 2>&1
tail -1 gpurun_out/r2m_plain2.log
