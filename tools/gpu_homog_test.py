"""How much of the rollout kernel's time is lane divergence and warp imbalance?  Same K, three noise batches:
  random      : K independent rows (the bench)
  warp_copies : every warp's 32 rollouts are copies of ONE row (lanes homogeneous, warps heterogeneous)
  all_copies  : all K rollouts are copies of one row (lanes homogeneous, warps balanced); averaged over a few rows
   python tools/gpu_homog_test.py <env> <force> <K>"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import engine as EN
env, force, K = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
eng = EN.RolloutEngine(env, device=0)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(12367134)
u = torch.zeros((2, 48), dtype=torch.float64, device=dev)
base = 0.8 * torch.randn((K, 2, 48), generator=gen, device=dev, dtype=torch.float32)
def t(noise):
    eng.physics(u, noise, [0, 0], force); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.physics(u, noise, [0, 0], force); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
out = {"env": env, "K": K, "random_ms": t(base)}
out["warp_copies_ms"] = t(base[::32].repeat_interleave(32, dim=0).contiguous())
out["all_copies_ms"] = [t(base[i:i + 1].expand(K, 2, 48).contiguous()) for i in (0, 1, 2, 3, 4, 5)]
print(json.dumps(out))
