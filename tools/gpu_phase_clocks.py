"""Phase breakdown of the rollout kernel (profiling builds compiled with -DPG_PHASE_CLOCKS, csrc/build_variant.sh): every warp's
lane 0 accumulates clock64() deltas around collide / external forces / solver set-up / Gauss-Seidel sweeps / integrate and adds
them to six global counters.  Prints the share of each phase in warp-cycles.
  PREGRASP_B200_LIB=.../libpregrasp_b200_clk*.so python tools/gpu_phase_clocks.py <env> <force> <K>"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import _abi, engine as EN          # noqa: E402

env, force, K = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
eng = EN.RolloutEngine(env, device=0)
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(12367134)
u = torch.zeros((2, 48), dtype=torch.float64, device=dev)
noise = 0.8 * torch.randn((K, 2, 48), generator=gen, device=dev, dtype=torch.float32)
pose = torch.empty((K, 2, 7), dtype=torch.float64, device=dev); fins = torch.empty((K, 2, 2, 3), dtype=torch.float64, device=dev)
metric = torch.empty(K, dtype=torch.float64, device=dev)
clk = torch.zeros(6, dtype=torch.int64, device=dev)
path = np.zeros(2, np.int32)
p = lambda t: C.c_void_p(t.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    clk.zero_()
    e0.record()
    _abi.check(eng.lib.pg_physics(eng.handle, p(u), p(noise), K, 2, path.ctypes.data_as(_abi._IP), force, p(pose), p(fins), p(metric), None, p(clk),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record(); torch.cuda.synchronize()
c = clk.cpu().numpy().astype(np.float64)
names = ["collide", "external", "solver_setup", "solver_sweeps", "integrate", "rest"]
print(json.dumps({"lib": os.path.basename(os.environ.get("PREGRASP_B200_LIB", "default")), "env": env, "K": K, "ms": round(e0.elapsed_time(e1), 1),
                  "share": {n: round(float(v / c.sum()), 4) for n, v in zip(names, c)}, "warp_Mcycles": {n: round(float(v / 1e6), 1) for n, v in zip(names, c)}}))
