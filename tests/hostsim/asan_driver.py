"""Runs the host-compiled rollout logic (AddressSanitizer + UBSan build) on a few random rollouts of every env.
Started by tests/test_oracle_cpu.py in a subprocess with libasan preloaded."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import _abi                      # noqa: E402
from synthesize_pregrasp_b200.envspec import make_spec         # noqa: E402

H = C.CDLL(sys.argv[1])
for env, F, K in (("plate", 20.0, 6), ("waterbottle", 20.0, 3), ("bookshelf", 50.0, 3), ("foodbox", 20.0, 3), ("laptop", 20.0, 2),
                  ("cardboard", 20.0, 2), ("keyboard", 20.0, 2)):
    spec = make_spec(env)
    cs, keep = _abi.make_c_spec(spec)
    N = 2
    rng = np.random.RandomState(1)
    noise = (0.8 * rng.randn(K, N, 48)).astype(np.float32)
    u = np.zeros((N, 48))
    pose = np.zeros((K, N, 7)); fins = np.zeros((K, N, 2, 3)); metric = np.zeros(K); path = np.zeros(N, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = H.hostsim_rollout(C.byref(cs), vp(u), vp(noise), K, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric), None, None)
    assert rc == 0 and np.isfinite(pose).all(), env
    print(env, "ok")
