#!/bin/bash
# Round 2, final evidence of the committed build: GPU tests, smoke, both bench arms, launch list of the default bench command,
# full-set ncu captures of rollout_kernel<box> (bench size, bench path -> profiles/ncu_traffic.json) and of pcn_tc_kernel.
# Every ncu command runs right after the same command exited 0 without ncu.
T=${1:-r2v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; cut -c1-300 gpurun_out/${T}_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; cut -c1-200 gpurun_out/${T}_bench_reference.json
CMD="python bench.py --steps 2 --warmup 3 --no-extra --cpu-sample 1"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bookshelf_K65536.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
wc -l gpurun_out/${T}_launches_bookshelf_K65536.csv
python tools/gpu_phys_time.py bookshelf 50 65536 1 0 bench > gpurun_out/${T}_plain2.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/${T}_rollout_box python tools/gpu_phys_time.py bookshelf 50 65536 1 0 bench > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_plain2.log gpurun_out/${T}_ncu2.log
python tools/gpu_cost_time.py bookshelf 16384 1 > gpurun_out/${T}_plain3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pcn_tc_kernel -s 1 -c 1 -o gpurun_out/${T}_cost python tools/gpu_cost_time.py bookshelf 16384 1 > gpurun_out/${T}_ncu3.log 2>&1
tail -2 gpurun_out/${T}_ncu3.log; ls -la gpurun_out | tail -8
timeout 300 python bench.py --workload mppi_update_K1048576 > gpurun_out/${T}_bench_mppi_update.json 2> gpurun_out/${T}_bench_mppi_update.err; cut -c1-400 gpurun_out/${T}_bench_mppi_update.json
