"""Kernel-LOGIC-vs-oracle agreement on the CPU: csrc/pg_rollout.cuh compiled for the host (tests/hostsim, no FMA contraction,
glibc math) against the independent oracle on seeded random rollouts -- separates algorithmic differences between the two
implementations from device effects (FMA contraction, CUDA libm).  Same statistics as tools/gpu_parity_report.py.
  python tools/host_parity_report.py [K] [env ...]"""
import ctypes as C, json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_spec
from synthesize_pregrasp_b200 import _abi
from synthesize_pregrasp_b200.envspec import make_spec
from oracle.envsim import EnvSim

K = int(sys.argv[1]) if len(sys.argv) > 1 else 48
envs = sys.argv[2:] or ["plate", "waterbottle", "foodbox", "bookshelf", "keyboard"]
FORCE = {"bookshelf": 50.0}
so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio", "-o", so,
                       os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
H = C.CDLL(so)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
out = {}
for env in envs:
    F = FORCE.get(env, 20.0)
    spec = make_spec(env); cs, keep = _abi.make_c_spec(spec)
    N = 2
    rng = np.random.RandomState(1234)
    noise = (0.8 * rng.randn(K, N, 48)).astype(np.float32)
    u = np.zeros((N, 48)); path = np.zeros(N, np.int32)
    pose = np.zeros((K, N, 7)); fins = np.zeros((K, N, 2, 3)); metric = np.zeros(K)
    assert H.hostsim_rollout(C.byref(cs), vp(u), vp(noise), K, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric), None, None) == 0
    errs = []
    for k in range(K):
        sim = EnvSim(oracle_spec(spec), F); sim.reset()
        ref = np.stack([sim.step(u[j] + noise[k, j].astype(np.float64), [0, 0, 0], train=True, n_steps=N)["final_pose"] for j in range(N)])
        errs.append(min(np.abs(pose[k] - ref).max(), np.abs(pose[k] * np.r_[1, 1, 1, -1, -1, -1, -1] - ref).max()))
    errs = np.array(errs)
    out[env] = dict(K=K, within_1e4=int((errs <= 1e-4).sum()), within_1e8=int((errs <= 1e-8).sum()), within_1e12=int((errs <= 1e-12).sum()),
                    identical=int((errs == 0).sum()), median=float(np.median(errs)), worst=[int(i) for i in np.argsort(-errs)[:6]],
                    worst_err=[float(e) for e in np.sort(errs)[::-1][:6]])
    print(env, json.dumps(out[env]), flush=True)
