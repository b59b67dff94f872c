#!/bin/bash
# Round 2, call p: whole GPU suite, smoke, default bench line with the plate extras, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2p_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2p_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_bench_default.json 2> gpurun_out/r2p_bench_default.err; tail -c 600 gpurun_out/r2p_bench_default.json; tail -3 gpurun_out/r2p_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; cut -c1-500 gpurun_out/r2p_bench_reference.json; tail -3 gpurun_out/r2p_bench_reference.err
