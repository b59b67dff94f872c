#!/bin/bash
# round-end evidence: GPU tests, smoke, default + plate bench, reference arm, ncu launch list of the default bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/final_pytest.log; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/final_bench_default.log 2>&1; tail -1 gpurun_out/final_bench_default.log | cut -c1-250
timeout 1500 python bench.py --steps 3 --warmup 3 --workload plate_K65536 > gpurun_out/final_bench_plate.log 2>&1; tail -1 gpurun_out/final_bench_plate.log | cut -c1-250
timeout 900 python bench.py --steps 3 --warmup 3 --workload plate_K360 > gpurun_out/final_bench_plate_K360.log 2>&1; tail -1 gpurun_out/final_bench_plate_K360.log | cut -c1-250
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.log 2>&1; tail -1 gpurun_out/final_bench_reference.log | cut -c1-400
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/final_ncu.log 2>&1; wc -l gpurun_out/final_launches.csv
