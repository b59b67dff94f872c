#!/bin/bash
# Round 2: GPU suite + smoke of the final library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1; tail -1 gpurun_out/r2t_smoke.log
