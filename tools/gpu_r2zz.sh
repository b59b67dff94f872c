#!/bin/bash
# Round 2, last build (double-precision GJK tolerances): GPU suite, smoke, hull-kernel physics time on the bench path
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s > gpurun_out/r2zz_pytest.log 2>&1; tail -3 gpurun_out/r2zz_pytest.log
grep "kernel vs PyBullet" gpurun_out/r2zz_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zz_smoke.log 2>&1; tail -1 gpurun_out/r2zz_smoke.log
timeout 100 python tools/gpu_phys_time.py plate 20 65536 2 0 bench > gpurun_out/r2zz_phys_plate.json 2>&1; cat gpurun_out/r2zz_phys_plate.json
timeout 60 python tools/gpu_phys_time.py plate 20 360 3 0 bench > gpurun_out/r2zz_phys_plate_K360.json 2>&1; cat gpurun_out/r2zz_phys_plate_K360.json
