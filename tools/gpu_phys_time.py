"""Times the physics stage (rollout kernel) of one workload and prints a digest of its outputs, so that kernel variants
(PREGRASP_B200_LIB=...) can be compared for speed AND bit-identity in one call.
  python tools/gpu_phys_time.py <env> <force> <K> [reps] [mode] [bench]
`bench`: the contact-graph path bench.py times (csg.npy / paths_id.npy[0]) instead of [0, 0] on the dummy graph."""
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
use_bench_path = len(sys.argv) > 6 and sys.argv[6] == "bench"
eng = EN.RolloutEngine(env, device=0, validate=use_bench_path)
path = [0, 0]
if use_bench_path:
    import bench                                             # noqa: E402
    path, on_csg = bench._bench_path(eng.spec)
    assert on_csg
eng.set_rollout_mode(mode)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(12367134)
u = torch.zeros((2, 48), dtype=torch.float64, device=dev)
noise = 0.8 * torch.randn((K, 2, 48), generator=gen, device=dev, dtype=torch.float32)
out = eng.physics(u, noise, path, force)
torch.cuda.synchronize()
ms = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = eng.physics(u, noise, path, force); e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
pose = out["pose"].cpu().numpy()
digest = hashlib.sha256(np.ascontiguousarray(pose).tobytes()).hexdigest()[:16]
print(json.dumps({"lib": os.path.basename(os.environ.get("PREGRASP_B200_LIB", "default")), "env": env, "K": K, "mode": mode, "path": [int(p) for p in path], "physics_ms": [round(m, 2) for m in ms],
                  "pose_sha": digest, "finite": bool(np.isfinite(pose).all())}))
