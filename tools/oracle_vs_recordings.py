"""Per-row error of the CPU oracle against every PyBullet recording held in tests/golden (the physics' only ground truth), with the
switches that explain three of them.  CPU only.
  python tools/oracle_vs_recordings.py > profiles/r2_oracle_vs_pybullet_recordings.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PHYS, oracle_spec, recorded, recording_time_spec   # noqa: E402
from oracle.envsim import EnvSim                                    # noqa: E402
from synthesize_pregrasp_b200.envspec import make_spec             # noqa: E402


def run(exp, params=None, committed_env=False):
    g = recorded(exp)
    osd = oracle_spec(make_spec(g["env"]) if committed_env else recording_time_spec(exp))
    sim = EnvSim(osd, g["max_force"], params=params)
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    return np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max(axis=1)


def summary(err):
    lead = lambda t: int(np.argmax(err > t)) if (err > t).any() else len(err)
    return {"rows_within": {"1e-12": lead(1e-12), "1e-10": lead(1e-10), "1e-8": lead(1e-8), "1e-6": lead(1e-6), "1e-4": lead(1e-4)},
            "max_over_100_rows": float(err.max()), "per_row": [float(f"{e:.2e}") for e in err]}


out = {"note": "max abs pose error per recorded row (q == -q); rows_within[t] = number of leading rows within t"}
for exp in PHYS:
    out[exp] = summary(run(exp))
out["wall_laptop_20.0 + parked thumb collides with the idle fingertips (tip_idle_collide=1)"] = summary(run("wall_laptop_20.0", dict(tip_idle_collide=1)))
out["keyboard_scale_20.0 with the COMMITTED dummy_states (thumb on region 9, where it touches; the recorded one, on region 2, hovers)"] = summary(run("keyboard_scale_20.0", committed_env=True))
out["cardboard_20.0 with the COMMITTED floor friction 0.002 (recording time: 0.005)"] = summary(run("cardboard_20.0", committed_env=True))
out["plate_20.0 with the single-precision GJK constants (REL_ERROR2 1e-6, penetration tolerance 1e-3, no GJK-normal check): the model before the last fix"] = summary(run("plate_20.0", dict(gjk_rel_error2=1e-6, gjk_pen_tol=1e-3, gjk_org_rule=0)))
out["cardboard2_20.0 (held out) with the thumb found before the floor in env step 2 (find_order_late=3429)"] = summary(run("cardboard2_20.0", dict(find_order_late=3429)))
out["cardboard3_20.0 (held out) with the thumb found before the floor in env step 2 (find_order_late=3429)"] = summary(run("cardboard3_20.0", dict(find_order_late=3429)))
out["plate_no_df_20.0 (held out) with the thumb found before the floor in env step 2 (find_order_late=3429)"] = summary(run("plate_no_df_20.0", dict(find_order_late=3429)))
out["plate_ablation_df_20.0 (held out) with the thumb found before the floor in env step 2 (find_order_late=3429)"] = summary(run("plate_ablation_df_20.0", dict(find_order_late=3429)))
out["round-1 model for comparison (no idle fingertips, round-1 pair order, no pre-refresh)"] = {
    exp: summary(run(exp, dict(idle_tips=0, find_order=12345, marginal_rule=1, tt_never_early=0, compound_prerefresh=0)))["rows_within"] for exp in PHYS}
print(json.dumps(out, indent=1))
