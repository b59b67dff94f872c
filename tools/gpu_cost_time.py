"""Times the cost stage (pcn_tc_kernel + mlp_kernel) on B random poses of one env.   python tools/gpu_cost_time.py <env> <B> [reps]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import engine as EN
env, B = sys.argv[1], int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
eng = EN.RolloutEngine(env, device=0)
spec = eng.spec
rng = np.random.RandomState(5)
pose = np.zeros((B, 7)); fins = np.zeros((B, 2, 3))
q = np.r_[spec.obj_quat] + 0.15 * rng.randn(B, 4); q /= np.linalg.norm(q, axis=1, keepdims=True)
pose[:, :3] = np.r_[spec.obj_pos] + 0.03 * rng.randn(B, 3) + [0, 0, 0.02]; pose[:, 3:] = q
fins[:] = spec.cloud[rng.randint(2048, size=(B, 2))]
pt, ft = torch.as_tensor(pose, device="cuda:0"), torch.as_tensor(fins, device="cuda:0")
eng.cost(pt, ft); torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); l = eng.cost(pt, ft); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print(json.dumps({"env": env, "B": B, "cost_ms": [round(m, 3) for m in ms], "us_per_eval_per_sm": round(min(ms) * 1e3 / (B / 148), 2), "logit_mean": float(l.nanmean())}))
