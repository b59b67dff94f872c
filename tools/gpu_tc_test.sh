#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -q -k "cost_kernel or score_entry" 2>&1 | tail -15 > gpurun_out/tc_test.log
cat gpurun_out/tc_test.log | tail -8
timeout 300 python - > gpurun_out/tc_time.log 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synthesize_pregrasp_b200 import engine as EN
from synthesize_pregrasp_b200.envspec import make_spec
spec = make_spec("bookshelf")
res = {}
for mode in ("tc", "ffma"):
    if mode == "ffma": os.environ["PG_COST_FFMA"] = "1"
    else: os.environ.pop("PG_COST_FFMA", None)
    eng = EN.RolloutEngine(spec, device=0)
    rng = np.random.RandomState(0)
    B = 16384
    pose = np.zeros((B, 7)); fins = np.zeros((B, 2, 3))
    for b in range(B):
        q = np.r_[spec.obj_quat] + 0.1 * rng.randn(4)
        pose[b] = np.r_[np.r_[spec.obj_pos] + 0.02 * rng.randn(3), q / np.linalg.norm(q)]
        fins[b] = spec.cloud[rng.randint(2048, size=2)]
    pt, ft = torch.as_tensor(pose, device="cuda:0"), torch.as_tensor(fins, device="cuda:0")
    out = eng.cost(pt, ft); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.cost(pt, ft); e1.record(); torch.cuda.synchronize()
    res[mode] = (e0.elapsed_time(e1), out.cpu().numpy())
    print(mode, "ms for", B, "evals:", res[mode][0])
d = np.abs(res["tc"][1] - res["ffma"][1])
print("max |tc - ffma| logit", np.nanmax(d), "median", np.nanmedian(d), "max |logit|", np.nanmax(np.abs(res["ffma"][1])))
PY
cat gpurun_out/tc_time.log | tail -5
