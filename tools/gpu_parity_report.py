"""Kernel-vs-oracle agreement report (GPU box): recorded controls (max trace difference over 100 substeps) and seeded
random batches (fraction of rollouts whose pose at both cost instants agrees within 1e-4 / 1e-8)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_spec, recorded
from synthesize_pregrasp_b200 import engine as EN
from synthesize_pregrasp_b200.envspec import make_spec
from oracle.envsim import EnvSim

def oracle_trace(spec, u, noise_row, F, train=False):
    sim = EnvSim(oracle_spec(spec), F); sim.reset(); rs = []
    for j in range(u.shape[0]):
        a = u[j] + (noise_row[j].astype(np.float64) if noise_row is not None else 0.0)
        rs.append(sim.step(a, [0, 0, 0], train=train, n_steps=u.shape[0]))
    return rs

out = {"recorded": {}, "random": {}}
for exp in ["plate_20.0", "waterbottle_20.0", "foodbox_20.0", "keyboard_20.0", "keyboard_scale_20.0", "wall_laptop_20.0", "cardboard_20.0"]:
    g = recorded(exp); spec = make_spec(g["env"]); eng = EN.RolloutEngine(spec, device=0)
    tr = eng.physics(g["u"], None, [0, 0], g["max_force"], K=1, want_trace=True)["trace"][0].cpu().numpy()
    P = np.concatenate([r["poses"] for r in oracle_trace(spec, g["u"], None, g["max_force"])])
    d = np.abs(tr - P).max(axis=1)
    out["recorded"][exp] = dict(max20=float(d[:20].max()), max50=float(d[:50].max()), max100=float(d.max()), identical_rows=int((d == 0).sum()))
K, N = int(os.environ.get("PR_K", 48)), 2
for env, F in [("plate", 20.0), ("waterbottle", 20.0), ("foodbox", 20.0), ("bookshelf", 50.0), ("keyboard", 20.0)]:
    spec = make_spec(env); eng = EN.RolloutEngine(spec, device=0)
    gen = torch.Generator(device="cuda:0").manual_seed(1234)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((N, 48))
    o = eng.physics(u, noise, [0, 0], F, want_trace=False)
    pose = o["pose"].cpu().numpy(); nz = noise.cpu().numpy()
    errs = []
    for k in range(K):
        ref = np.stack([r["final_pose"] for r in oracle_trace(spec, u, nz[k], F, train=True)])
        errs.append(min(np.abs(pose[k] - ref).max(), np.abs(pose[k] * np.r_[1, 1, 1, -1, -1, -1, -1] - ref).max()))
    errs = np.array(errs)
    out["random"][env] = dict(K=K, within_1e4=int((errs <= 1e-4).sum()), within_1e8=int((errs <= 1e-8).sum()), within_1e12=int((errs <= 1e-12).sum()), median=float(np.median(errs)))
print(json.dumps(out, indent=1))
