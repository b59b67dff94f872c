#!/bin/bash
# Round 2, last test state (keyboard_scale with its recording-time contact state, whole-horizon bounds): GPU suite + smoke
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -s > gpurun_out/r2zz2_pytest.log 2>&1; tail -3 gpurun_out/r2zz2_pytest.log
grep "kernel vs PyBullet" gpurun_out/r2zz2_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2zz2_smoke.log 2>&1; tail -1 gpurun_out/r2zz2_smoke.log
