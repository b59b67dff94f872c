"""Whole-rollout oracle: decode -> physics -> cost, for ONE sample (test infrastructure).
Follows the worker loop stoch_traj_opt.py:124-136: cost = sum_j -(100 * score logit) over the env steps."""
import numpy as np

from . import cost as CO
from . import scorenet as SN
from .envsim import EnvSim


def rollout(spec, cost_spec, weights, u, noise_row, path, max_force, params=None, net_dtype=np.float64):
    """spec: EnvSim dict; cost_spec: dict(cloud, deocclude_boxes, dist_boxes, dist_tris, clearance_h,
    walls=[(pos,quat)], wall_size, strict_z); u [N,48]; noise_row [N,48] or None.
    -> dict(J, metric, poses [N,7], fins [N,2,3], logits [N], trace [N*50,7])"""
    N = u.shape[0]
    sim = EnvSim(spec, max_force, params)
    sim.reset()
    J, metric = 0.0, 0.0
    poses, fins, logits, trace = [], [], [], []
    for j in range(N):
        a = np.asarray(u[j], np.float64) + (np.asarray(noise_row[j], np.float32).astype(np.float64) if noise_row is not None else 0.0)
        r = sim.step(a, path, train=True, n_steps=N)
        pose = r["final_pose"]
        pts, df, _ = CO.cost_inputs(pose[:3], pose[3:], cost_spec["cloud"], cost_spec["deocclude_boxes"],
                                    cost_spec["dist_boxes"], cost_spec.get("dist_tris"))
        grasp = np.concatenate([CO.project_to_pcd(r["cur_pos"][f], cost_spec["cloud"], cost_spec["clearance_h"],
                                                  cost_spec["walls"], cost_spec["wall_size"], cost_spec["strict_z"])
                                for f in range(2)])
        logit = SN.score_logit(weights, pts, df, grasp, net_dtype) if len(pts) else float("nan")
        J += np.inf if np.isnan(logit) else -100.0 * logit
        metric += r["metric"]
        poses.append(pose); fins.append(r["cur_pos"]); logits.append(logit); trace.append(r["poses"])
    sim.close()
    return dict(J=J, metric=metric, poses=np.array(poses), fins=np.array(fins), logits=np.array(logits),
                trace=np.concatenate(trace))
