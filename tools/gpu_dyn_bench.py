"""Throughput of the feasibility sweeps: all C(P,3) triples of a box-surface cloud, simple and QP criteria."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import dyn_feasibility_check as DYN

def box_cloud(P, seed=0, size=(0.4, 0.4, 0.1)):          # dyn_feasibility_check.py:312-317 uses a poisson sample of this box
    rng = np.random.RandomState(seed)
    h = np.array(size) / 2
    pts, nrm = [], []
    area = np.array([size[1] * size[2], size[0] * size[2], size[0] * size[1]])
    for _ in range(P):
        ax = rng.choice(3, p=area / area.sum()); sgn = rng.choice([-1.0, 1.0])
        p = rng.uniform(-h, h); p[ax] = sgn * h[ax]
        n = np.zeros(3); n[ax] = sgn
        pts.append(p); nrm.append(n)
    return np.array(pts), np.array(nrm)

out = {}
for P in (256, 512):
    pts, nrm = box_cloud(P)
    total = P * (P - 1) * (P - 2) // 6
    for crit in ("simple", "qp"):
        DYN.find_dynamically_feasible_contacts(pts, nrm, criterion=crit, max_out=1 << 20)      # warm-up
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tri, n = DYN.find_dynamically_feasible_contacts(pts, nrm, criterion=crit, max_out=1 << 20)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        out[f"P{P}_{crit}"] = dict(triples=total, feasible=n, seconds=dt, triples_per_s=total / dt)
print(json.dumps(out, indent=1))
