"""Generate tests/golden/*.npz by running the REFERENCE's own Python where it is runnable here.

Build container only (needs /root/reference).  What runs is the reference's unmodified source,
imported with stub modules for the third-party packages that are absent from this image
(open3d, cvxpy, pytorch_kinematics, pointnet2_ops, pybullet, geomdl, tqdm are stubbed ONLY as far as
import-time symbols go; the two pytorch_kinematics functions the feasibility test actually calls are
re-implemented from their published definitions).
  score_net.npz      neurals/network.py::LargeScoreFunction.pred_score on CPU, fp32 and fp64
  dyn_simple.npz     utils/dyn_feasibility_check.py::simple_check_dyn_feasible(_3p)
  mppi_update.npz    the update lines of stoch_traj_opt.py:195-204 executed verbatim
  recorded_*.npz     the reference's recorded (u -> object pose, tip pose) trajectories (data/traj,
                     data/object_poses, data/tip_data), first 100 rows = the train-mode control substeps
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PREGRASP_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, ".."))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k: (_ for _ in ()).throw(AttributeError(k)) if k.startswith("__") else types.SimpleNamespace()
    sys.modules[name] = m
    return m


def _pk_axis_angle_to_quaternion(axis_angle):
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    eps = 1e-6
    small = angles.abs() < eps
    s = torch.empty_like(angles)
    s[~small] = torch.sin(half[~small]) / angles[~small]
    s[small] = 0.5 - (angles[small] * angles[small]) / 48
    return torch.cat([torch.cos(half), axis_angle * s], dim=-1)


def _pk_raw_mul(a, b):
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def _pk_quaternion_apply(q, p):
    pq = torch.cat((p.new_zeros(p.shape[:-1] + (1,)), p), -1)
    inv = q * q.new_tensor([1, -1, -1, -1])
    return _pk_raw_mul(_pk_raw_mul(q, pq), inv)[..., 1:]


def install_stubs():
    for n in ("open3d", "cvxpy", "pybullet", "geomdl", "pygeodesic", "pygeodesic.geodesic"):
        _stub(n)
    _stub("pytorch_kinematics", axis_angle_to_quaternion=_pk_axis_angle_to_quaternion,
          quaternion_apply=_pk_quaternion_apply)
    _stub("pointnet2_ops")
    _stub("pointnet2_ops.pointnet2_modules")
    try:
        import tqdm  # noqa
    except ImportError:
        _stub("tqdm", tqdm=lambda x, **k: x)


def gen_score(rng):
    from neurals.network import LargeScoreFunction
    net = LargeScoreFunction(num_fingers=2, latent_dim=20, has_distance_field=True)
    net.load_state_dict(torch.load(os.path.join(
        REF, "neurals/pretrained_score_function/only_score_model_with_df/2980.pth"), map_location="cpu"))
    net.eval()
    plate = np.load(os.path.join(HERE, "..", "synthesize_pregrasp_b200", "assets", "plate.npz"))
    cases = {}
    for i, P in enumerate([2048, 1500, 977, 1]):
        idx = rng.permutation(2048)[:P]
        pts = plate["cloud"][idx].astype(np.float32)
        df = np.abs(rng.normal(0.05, 0.03, P)).astype(np.float32)
        grasp = np.concatenate([pts[rng.randint(P)], [100, 100, 100] if i == 2 else pts[rng.randint(P)]]).astype(np.float32)
        with torch.no_grad():
            s32 = net.pred_score(torch.from_numpy(pts)[None], torch.from_numpy(grasp)[None],
                                 torch.from_numpy(df)[None, :, None]).item()
            s64 = net.double().pred_score(torch.from_numpy(pts)[None].double(), torch.from_numpy(grasp)[None].double(),
                                          torch.from_numpy(df)[None, :, None].double()).item()
            net.float()
        cases[f"pts{i}"], cases[f"df{i}"], cases[f"grasp{i}"] = pts, df, grasp
        cases[f"logit32_{i}"], cases[f"logit64_{i}"] = np.float64(s32), np.float64(s64)
    np.savez_compressed(os.path.join(OUT, "score_net.npz"), n=4, **cases)
    print("score", [cases[f"logit64_{i}"] for i in range(4)])


def gen_dyn(rng):
    import utils.dyn_feasibility_check as ref
    pos = np.load(os.path.join(REF, "data/dyn_feasible/0.npy"))[:: 97][:200]          # recorded positives
    pts, nrm, out = [], [], []
    for row in pos:
        pts.append(row[:, :3]); nrm.append(row[:, 3:]); out.append(ref.simple_check_dyn_feasible(row[:, :3], row[:, 3:]))
    box = np.array([0.2, 0.2, 0.05])
    for _ in range(1200):                                                               # random box-surface triples
        p = np.zeros((3, 3)); n = np.zeros((3, 3))
        for i in range(3):
            ax = rng.randint(3); sg = rng.choice([-1.0, 1.0])
            q = rng.uniform(-1, 1, 3) * box; q[ax] = sg * box[ax]
            p[i] = q; n[i, ax] = -sg
            if rng.rand() < 0.3:
                n[i] += rng.normal(0, 0.3, 3); n[i] /= np.linalg.norm(n[i])
        pts.append(p); nrm.append(n); out.append(ref.simple_check_dyn_feasible(p, n))
    m_pts, m_nrm, m_out = [], [], []
    for _ in range(100):                                                                # M=4/5 contacts
        M = rng.choice([4, 5]); p = rng.uniform(-1, 1, (5, 3)) * box; n = rng.normal(0, 1, (5, 3))
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        p[M:] = 0; n[M:] = 0
        m_pts.append(p); m_nrm.append(n); m_out.append(ref.simple_check_dyn_feasible(p[:M], n[:M]))
    np.savez_compressed(os.path.join(OUT, "dyn_simple.npz"), pts=np.array(pts), nrm=np.array(nrm),
                        out=np.array(out, bool), m_pts=np.array(m_pts), m_nrm=np.array(m_nrm),
                        m_out=np.array(m_out, bool), m_count=np.array([int((np.abs(p).sum(1) > 0).sum()) for p in m_pts]))
    print("dyn: positives", int(np.sum(out)), "of", len(out), "| multi", int(np.sum(m_out)), "of", len(m_out))


def gen_mppi(rng):
    K, N, D, kappa, alpha = 360, 2, 48, 5, 1
    u = rng.normal(0, 0.3, (N, D)); E = 0.8 * rng.normal(0, 1, (N, D, K)); J = rng.normal(700, 30, K)
    u0 = u.copy()
    # ---- stoch_traj_opt.py:195-204, verbatim ----
    J2 = J - min(J)
    S = np.exp(-kappa * J2)
    norm = np.sum(S)
    S = S / norm
    for j in range(N):
        for k in range(D):
            u[j, k] = u[j, k] + alpha * np.sum(S * E[j, k, :])
    np.savez_compressed(os.path.join(OUT, "mppi_update.npz"), u0=u0, J=J, E=E, u1=u, S=S, kappa=kappa, alpha=alpha)


def gen_recorded():
    exps = {"plate_20.0": "plate", "foodbox_20.0": "foodbox", "foodbox_50.0": "foodbox", "keyboard_20.0": "keyboard",
            "keyboard_50.0": "keyboard", "keyboard_scale_20.0": "keyboard", "wall_laptop_20.0": "laptop",
            "cardboard_20.0": "cardboard", "waterbottle_20.0": "waterbottle", "book_shelf_50.0": "bookshelf",
            # held out until the physics model was frozen (end of round 2), then replayed without touching it:
            "cardboard_scale_20.0": "cardboard", "cardboard2_20.0": "cardboard", "cardboard3_20.0": "cardboard",
            "foodbox_ablation_df_20.0": "foodbox", "plate_ablation_df_20.0": "plate", "plate_no_df_20.0": "plate"}
    out = {}
    for exp, env in exps.items():
        k = exp.replace(".", "p")
        out[k + "__u"] = np.load(os.path.join(REF, f"data/traj/{exp}.npy"))
        out[k + "__poses"] = np.load(os.path.join(REF, f"data/object_poses/{exp}_object_poses.npy"))[:100]
        out[k + "__tips"] = np.load(os.path.join(REF, f"data/tip_data/{exp}_tip_poses.npy"))[:100, :2, :3]
        out[k + "__env"] = np.array(env)
        out[k + "__max_force"] = np.float64(exp.split("_")[-1])
    np.savez_compressed(os.path.join(OUT, "recorded_trajectories.npz"), **out)
    # full fingertip records (position + normal, all five slots) of the recordings whose normals the stored clouds reproduce:
    # the target of synthesize_pregrasp_b200/model_test.py::tip_records
    full = {}
    for exp in ("plate_20.0", "foodbox_20.0", "waterbottle_20.0", "wall_laptop_20.0"):
        full[exp.replace(".", "p")] = np.load(os.path.join(REF, f"data/tip_data/{exp}_tip_poses.npy"))[:100]
    np.savez_compressed(os.path.join(OUT, "recorded_tip_records.npz"), **full)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    install_stubs()
    rng = np.random.RandomState(20231019)
    gen_recorded()
    gen_mppi(rng)
    gen_score(rng)
    gen_dyn(rng)
    print(sorted(os.listdir(OUT)))
