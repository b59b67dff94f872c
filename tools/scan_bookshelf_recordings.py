"""Which env was each of the reference's bookshelf recordings made with?  (VERDICT r1 item 3 / SURVEY 8(c).)
Build container only (reads /root/reference/data).  For every data/traj/*book*shelf*.npy with a matching object-pose /
fingertip recording the script derives, FROM THE DATA: which fingers are on in each env step, the local contact point of every
active finger (object frame: recorded tip position pulled back through the recorded object pose), the face offset that point
implies (local coordinate minus the 0.01 stand-off of plate_env:279), where parked fingers sit, and the object pose after the
reset + pre-simulate substeps (row 0).  It then runs the oracle of the COMMITTED env (box 0.4 x 0.4 x 0.05, SCALE = (1, 1, 0.5),
envs/scales.py) on the recorded controls and reports the row-0 / whole-horizon error.
  python tools/scan_bookshelf_recordings.py > profiles/r2_bookshelf_recordings.json"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_spec                                   # noqa: E402
from oracle.envsim import EnvSim                                  # noqa: E402
from synthesize_pregrasp_b200.envspec import make_spec            # noqa: E402

R = os.environ.get("PREGRASP_REFERENCE", "/root/reference") + "/data"


def qrot_inv(q, v):
    x, y, z, w = q
    u = -np.array([x, y, z])
    return v + 2 * np.cross(u, np.cross(u, v) + w * v)


out = {}
for f in sorted(glob.glob(R + "/traj/*ook*shelf*.npy")):
    exp = os.path.basename(f)[:-4]
    u = np.load(f)
    try:
        poses = np.load(f"{R}/object_poses/{exp}_object_poses.npy")[:100]
        tips = np.load(f"{R}/tip_data/{exp}_tip_poses.npy")[:100]
    except FileNotFoundError:
        out[exp] = {"usable": False, "why": "no object_poses / tip_data recording"}
        continue
    force = float(exp.split("_")[-1])
    rec = {"max_force": force, "row0_pose": [float(x) for x in poses[0]], "steps": []}
    for j in range(2):
        st = {}
        for fi in range(2):
            a = np.tanh(u[j, fi * 24:(fi + 1) * 24])
            loc = qrot_inv(poses[j * 50, 3:], tips[j * 50, fi, :3] - poses[j * 50, :3])
            on = bool(np.abs(loc).max() < 10)
            st[f"finger{fi}"] = {"on": on, "local": [round(float(x), 5) for x in loc], "on_off_action": round(float(a[23]), 3)}
            if on:                                   # which face: the coordinate closest to (half extent + 0.01 stand-off)
                st[f"finger{fi}"]["face_offsets_minus_standoff"] = [round(abs(float(x)) - 0.01, 5) for x in loc]
        rec["steps"].append(st)
    both_off = all(not s[f"finger{fi}"]["on"] for s in rec["steps"] for fi in range(2))
    parked = [s[f"finger{fi}"]["local"] for s in rec["steps"] for fi in range(2) if not s[f"finger{fi}"]["on"]]
    rec["parked_local"] = parked[0] if parked else None
    # committed env on the recorded controls
    spec = make_spec("bookshelf")
    sim = EnvSim(oracle_spec(spec), force)
    sim.reset()
    P = np.concatenate([sim.step(u[j], [0, 0, 0])["poses"] for j in range(2)])
    err = np.minimum(np.abs(P - poses), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - poses)).max(axis=1)
    rec["committed_env"] = {"row0_err": float(err[0]), "max_err_100_rows": float(err.max()), "oracle_row0_pose": [float(x) for x in P[0]],
                            "oracle_row99_pose": [float(x) for x in P[99]], "recorded_row99_pose": [float(x) for x in poses[99]]}
    zs = [s[f"finger{fi}"]["local"][2] for s in rec["steps"] for fi in range(2) if s[f"finger{fi}"]["on"] and abs(abs(s[f"finger{fi}"]["local"][2]) - 0.035) > 1e-3
          and abs(s[f"finger{fi}"]["local"][0]) < 0.205 and abs(s[f"finger{fi}"]["local"][1]) < 0.205]
    verdict = []
    if parked and abs(parked[0][2] - 50.0) < 1e-3:
        verdict.append("parked fingers at local z = 50 = 100 x SCALE_z: an older env multiplied the parking position by SCALE (the committed env parks at (100,100,100))")
    tops = sorted({round(abs(z), 4) for z in zs})
    if tops:
        verdict.append(f"top/bottom-face contacts at |local z| = {tops} instead of the committed 0.025 + 0.01 = 0.035: book thickness (SCALE_z) differed")
    if both_off:
        verdict.append("both fingers off in both env steps: object-only physics; row 0 differs from the committed box env in the 5th digit -- "
                       "the magnitude of a hull-vs-floor first contact (book.urdf's commented-out <mesh filename=\"box.stl\">), not of a box-box 4-point contact")
    rec["what_drifted"] = verdict
    out[exp] = rec
print(json.dumps(out, indent=1))
