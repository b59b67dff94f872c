"""FLOP audit of the rollout kernel's Gauss-Seidel sweep, per physics substep (SURVEY 8(d) asks for the estimate to be replaced
by a count).  Compiles the kernel's per-thread code for the host with -DPG_HOST_STATS (tests/hostsim/hostsim.cpp; the counters
are host-only and compile to nothing in the CUDA build), runs seeded random rollouts of the bench workloads and prints the
audited FLOPs (1 FMA = 2) of the solver rows: servo rows, normal rows, friction pairs, and the mean iteration count.
Collision detection, integration and decode are not counted (they are outside the iteration loop).

  python tools/audit_substep_flops.py [rollouts per env]        ->  JSON on stdout (committed as profiles/r2_flop_audit.json; the bench workloads: csg.npy + paths_id.npy[0])
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import _abi                      # noqa: E402
from synthesize_pregrasp_b200.envspec import make_spec         # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim_stats.so")
subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-DPG_HOST_STATS", "-include", "cstdio",
                       "-o", so, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
H = C.CDLL(so)
out = {}
for env, F in (("bookshelf", 50.0), ("plate", 20.0), ("waterbottle", 20.0), ("foodbox", 20.0), ("keyboard", 20.0), ("laptop", 20.0), ("cardboard", 20.0)):
    spec = make_spec(env)
    # the bench's contact graph and path (bench.py _bench_path): csg.npy + paths_id.npy[0] where the reference ships them
    validate = spec.paths_id is not None and spec.csg_states is not None
    cs, keep = _abi.make_c_spec(spec, spec.csg_states if validate else None)
    N = 2
    rng = np.random.RandomState(12367134 % (2 ** 31))
    noise = (0.8 * rng.randn(K, N, 48)).astype(np.float32)
    u = np.zeros((N, 48))
    pose = np.zeros((K, N, 7)); fins = np.zeros((K, N, 2, 3)); metric = np.zeros(K)
    path = np.asarray(spec.paths_id[0][:N] if validate else [0] * N, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    fl = np.zeros(4)
    H.hostsim_flops(vp(fl), 1)
    assert H.hostsim_rollout(C.byref(cs), vp(u), vp(noise), K, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric), None, None) == 0
    H.hostsim_flops(vp(fl), 1)
    substeps = K * spec.substeps_per_rollout
    out[env] = {"rollouts": K, "path": [int(x) for x in path], "contact_graph": "csg.npy" if validate else "dummy_states.npy", "substeps_per_rollout": spec.substeps_per_rollout,
                "solver_flop_per_substep": float(fl[:3].sum() / substeps),
                "servo_rows": float(fl[0] / substeps), "normal_rows": float(fl[1] / substeps), "friction_pairs": float(fl[2] / substeps),
                "iterations_per_substep": float(fl[3] / substeps)}
print(json.dumps(out, indent=1))
