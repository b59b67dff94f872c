"""Times the physics stage (rollout kernel) of one workload and prints a digest of its outputs, so that kernel variants
(PREGRASP_B200_LIB=...) can be compared for speed AND bit-identity in one call.
  python tools/gpu_phys_time.py <env> <force> <K> [reps] [mode]"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import engine as EN          # noqa: E402

env, force, K = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 0          # pg_set_rollout_mode: 0 auto, 1 persistent launch, 2 phased + work-sorted
eng = EN.RolloutEngine(env, device=0)
eng.set_rollout_mode(mode)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(12367134)
u = torch.zeros((2, 48), dtype=torch.float64, device=dev)
noise = 0.8 * torch.randn((K, 2, 48), generator=gen, device=dev, dtype=torch.float32)
out = eng.physics(u, noise, [0, 0], force)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.physics(u, noise, [0, 0], force); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
pose = out["pose"].cpu().numpy()
digest = hashlib.sha256(np.ascontiguousarray(pose).tobytes()).hexdigest()[:16]
print(json.dumps({"lib": os.path.basename(os.environ.get("PREGRASP_B200_LIB", "default")), "env": env, "K": K, "mode": mode, "physics_ms": [round(m, 2) for m in ms],
                  "pose_sha": digest, "finite": bool(np.isfinite(pose).all())}))
