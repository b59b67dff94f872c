#!/bin/bash
# Round 2, call x: whole GPU suite (no -x) on the build with the idle fingertips in the dispatcher order
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/r2x_pytest.log 2>&1; grep -E "PASSED|FAILED|passed|failed|kernel vs PyBullet|within" gpurun_out/r2x_pytest.log | tail -70
