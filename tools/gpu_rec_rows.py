"""Per-row error of the CUDA kernel on one recorded trajectory: against the PyBullet recording and against the oracle.
  python tools/gpu_rec_rows.py <exp> [<exp> ...]      (PREGRASP_B200_LIB selects a library variant)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PHYS, recorded                                  # noqa: E402
from test_gpu_parity import _oracle_trace, _recording_time_spec     # noqa: E402
from synthesize_pregrasp_b200 import engine as EN                   # noqa: E402

for exp in sys.argv[1:]:
    g = recorded(exp)
    spec = _recording_time_spec(exp, g["env"])
    eng = EN.RolloutEngine(spec, device=0)
    tr = eng.physics(g["u"], None, [0, 0], g["max_force"], K=1, want_trace=True)["trace"][0].cpu().numpy()
    P = np.concatenate([r["poses"] for r in _oracle_trace(spec, g["u"], None, g["max_force"])])
    d = np.abs(tr - P).max(axis=1)
    e = np.minimum(np.abs(tr - g["poses"]), np.abs(tr * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max(axis=1)
    lead = lambda x, t: int(np.argmax(x > t)) if (x > t).any() else len(x)
    print(os.path.basename(os.environ.get("PREGRASP_B200_LIB", "default")), exp, "vs recording: rows within 1e-10 / 1e-8 / 1e-6 / 1e-4:",
          lead(e, 1e-10), lead(e, 1e-8), lead(e, 1e-6), lead(e, 1e-4), "max %.1e" % e.max(),
          "| vs oracle: rows within 1e-10 / 1e-6:", lead(d, 1e-10), lead(d, 1e-6), "max %.1e" % d.max())
