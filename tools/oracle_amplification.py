"""How fast does each recorded trajectory amplify a rounding-level difference?  The oracle is run twice per recording, the second
time with the object's velocity nudged by 1e-12 at one early world step (oracle knob `perturb`), and the distance between the two
runs is printed next to the oracle's error against the PyBullet recording.  Where the two curves have the same shape the error is
what ANY implementation that does not reproduce PyBullet's binary bit for bit must show (foodbox_20.0: x150 and x100 in world steps
15 and 16, x30 per step at rows 29-31); where the error jumps and the nudge does not, a rule is still missing (waterbottle_20.0 row
37, wall_laptop_20.0 row 7).  CPU only.
  python tools/oracle_amplification.py > profiles/r2_amplification.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PHYS, oracle_spec, recorded, recording_time_spec   # noqa: E402
from oracle.envsim import EnvSim                                        # noqa: E402


def run(exp, params=None):
    g = recorded(exp)
    sim = EnvSim(oracle_spec(recording_time_spec(exp)), g["max_force"], params=params)
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    return P, np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max(axis=1)


out = {"note": "per recorded row: `error` = oracle vs PyBullet recording; `nudged` = oracle vs the same oracle with the object's velocity "
               "changed by 1e-12 at world step `nudge_step`; `gain` = largest one-row growth factor of `nudged`"}
for exp in PHYS:
    P0, err = run(exp)
    P1, _ = run(exp, dict(perturb=1e-12, dbg_step=6))
    d = np.abs(P1 - P0).max(axis=1)
    nz = d > 0
    gain = float(np.max(d[1:][nz[:-1]] / d[:-1][nz[:-1]])) if nz[:-1].any() else 0.0
    out[exp] = {"nudge_step": 6, "final_nudged": float(d[-1]), "final_error": float(err[-1]), "gain": round(gain, 1),
                "error": [float(f"{e:.1e}") for e in err], "nudged": [float(f"{e:.1e}") for e in d]}
print(json.dumps(out, indent=1))
