#!/bin/bash
# quick device check of a build: smoke, recorded trajectories per row, physics time + digest
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2q2.log
timeout 300 python tools/gpu_rec_rows.py cardboard_20.0 keyboard_20.0 plate_20.0 foodbox_20.0 2>&1 | grep "vs recording" | tee -a gpurun_out/r2q2.log
timeout 300 python tools/gpu_phys_time.py bookshelf 50 65536 2 0 bench 2>&1 | tail -1 | tee -a gpurun_out/r2q2.log
