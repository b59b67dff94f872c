"""Host-side mirror of the reference's replay / recording script, model_test.py:115-135 (without --add_physics): replay
optimised controls `u[N,48]` once with `train=False` and record, for every control substep, the object pose and the five
fingertip records the later pipeline stages read:

  data/object_poses/<exp>_object_poses.npy   [N*50, 7]      pos xyz + quat xyzw at the START of each control substep
  data/tip_data/<exp>_tip_poses.npy          [N*50, 5, 6]   per finger: world contact point (object pose o local point,
                                                            plate_env:513-520) and world normal of the nearest cloud point
                                                            (`compute_tip_normals`, :676-679); absent fingers = 100 (`parse_tip_data`, :28-33)

The poses come from ONE K=1 launch of the rollout kernel with the trace output (`pg_physics(..., trace_dev)`), the fingertip
records are derived from them and from the decoded local contact points on the host.  Not mirrored: rendering, and the
`--add_physics` grasp-and-lift continuation (rows >= N*50 of the 200-row recordings; SURVEY.md 8f N1: three more constrained
tips and a mass change).  Normals: the nearest-neighbour normal of the env's stored 2048-point cloud reproduces the recorded
ones to 2e-7 for plate, foodbox, waterbottle and laptop; the keyboard / cardboard recordings were made with clouds whose
normals follow another convention, so for those envs only the positions are comparable.

    python -m synthesize_pregrasp_b200.model_test --exp_name plate_20.0 --env plate          # reads data/traj/plate_20.0.npy
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import envspec as ES

NUM_TIP_SLOTS = 5                    # model_test.py:29


def multiply_transforms(posA, ornA, posB):
    """pybullet.multiplyTransforms(posA, ornA, posB, [0,0,0,1])[0] as PyBullet computes it: in SINGLE precision (b3Transform with
    b3Scalar = float), every operation a float32 operation in b3's order.  The reference's recorded fingertip files hold exactly
    these float32 values (reproduced bit for bit from the recorded poses by tests/test_oracle_cpu.py); the rollout kernel computes
    its servo targets the same way (csrc/pg_math.cuh b3_multiply_transforms)."""
    f = np.float32
    x, y, z, w = (f(v) for v in ornA)
    d = f(f(f(f(x * x) + f(y * y)) + f(z * z)) + f(w * w))
    s = f(f(2.0) / d)
    xs, ys, zs = f(x * s), f(y * s), f(z * s)
    wx, wy, wz = f(w * xs), f(w * ys), f(w * zs)
    xx, xy, xz = f(x * xs), f(x * ys), f(x * zs)
    yy, yz, zz = f(y * ys), f(y * zs), f(z * zs)
    one = f(1.0)
    R = ((f(one - f(yy + zz)), f(xy - wz), f(xz + wy)),
         (f(xy + wz), f(one - f(xx + zz)), f(yz - wx)),
         (f(xz - wy), f(yz + wx), f(one - f(xx + yy))))
    v = [f(t) for t in posB]
    o = [f(t) for t in posA]
    out = np.empty(3)
    for i in range(3):
        acc = f(f(f(v[0] * R[i][0]) + f(v[1] * R[i][1])) + f(v[2] * R[i][2]))
        out[i] = f(acc + o[i])
    return out


def tip_records(object_poses, local_points, cloud, cloud_normals, control_skip=ES.CONTROL_SKIP):
    """object_poses [N*skip,7], local_points [N,F,3] (cur_fins of each env step; (100,100,100) for a finger that is off)
    -> tip_poses [N*skip, 5, 6] exactly as `parse_tip_data(info["finger_pos"])` lays them out."""
    object_poses = np.asarray(object_poses, np.float64)
    local_points = np.asarray(local_points, np.float64)
    T, F = object_poses.shape[0], local_points.shape[1]
    out = np.full((T, NUM_TIP_SLOTS, 6), 100.0)
    for j in range(local_points.shape[0]):
        # compute_tip_normals: nearest neighbour of the LOCAL point in the object's cloud (KD-tree query, k = 1)
        nn = [int(np.argmin(((cloud - local_points[j, f]) ** 2).sum(axis=1))) for f in range(F)]
        for t in range(control_skip):
            row = j * control_skip + t
            if row >= T:
                break
            pos, quat = object_poses[row, :3], object_poses[row, 3:]
            for f in range(F):
                out[row, f, :3] = multiply_transforms(pos, quat, local_points[j, f])              # plate_env:515
                out[row, f, 3:] = multiply_transforms(np.zeros(3), quat, cloud_normals[nn[f]])    # plate_env:518
    return out


def record(env_name, u, max_force, path=None, validate=False, device=0):
    """-> (tip_poses [N*50,5,6], object_poses [N*50,7], J, metric) of one replay of `u` (model_test.py:115-130)."""
    from . import engine as EN
    spec = ES.make_spec({"wallbox": "laptop"}.get(env_name, env_name))
    eng = EN.RolloutEngine(spec, device=device, validate=validate)
    u = np.asarray(u, np.float64)
    N = u.shape[0]
    path = [0] * N if path is None else list(path)
    out = eng.physics(u, None, path, max_force, K=1, want_trace=True)
    object_poses = out["trace"][0].cpu().numpy()
    fins = out["fins"][0].cpu().numpy()
    J, metric = eng.rollout(u, None, path, max_force, K=1, want_metric=True)
    tips = tip_records(object_poses, fins, spec.cloud, spec.cloud_normals)
    eng.close()
    return tips, object_poses, float(J.item()), float(metric.item())


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--exp_name", type=str, required=True)
    p.add_argument("--env", type=str, default="foodbox")
    p.add_argument("--validate", action="store_true", default=False)
    p.add_argument("--path_id", type=int, default=0)
    p.add_argument("--data_dir", type=str, default="data")
    a = p.parse_args(argv)
    u = np.load(os.path.join(a.data_dir, "traj", f"{a.exp_name}.npy"))
    max_force = float(a.exp_name.split("_")[-1])               # model_test.py:65-68
    if max_force == 0:
        max_force = 80.0
    path = None
    if a.validate:                                             # model_test.py:82-85
        spec = ES.make_spec({"wallbox": "laptop"}.get(a.env, a.env))
        path = [int(s) for s in np.asarray(spec.paths_id)[a.path_id]]
    tips, poses, J, metric = record(a.env, u, max_force, path, a.validate)
    for sub in ("tip_data", "object_poses"):
        os.makedirs(os.path.join(a.data_dir, sub), exist_ok=True)
    np.save(os.path.join(a.data_dir, "tip_data", f"{a.exp_name}_tip_poses.npy"), tips)
    np.save(os.path.join(a.data_dir, "object_poses", f"{a.exp_name}_object_poses.npy"), poses)
    print(f"reward {-J:.4f} metric {metric:.4f}; {poses.shape[0]} substeps recorded")


if __name__ == "__main__":
    main()
