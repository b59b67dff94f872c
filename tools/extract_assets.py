"""Convert the reference's DATA assets (no code) into the compact bundles the product loads.

Run in the build container only (needs /root/reference):  python tools/extract_assets.py
Outputs  synthesize_pregrasp_b200/assets/*.npz  (committed; a few hundred kB).

Sources (all relative to the reference root):
  envs/assets/plate.ply                         object cloud of the plate env (plate_contact_graph_env.py:202-205)
  envs/assets/plate_cvx_simple.obj              collision hull + MeshRegion triangles (plate_contact_graph_env.py:98-101)
  data/regions/small_block_dummy_region.npz     SmallBlockRegionDummy tables (utils/small_block_region.py:8-16)
  data/contact_states/<env>_env/*.npy           contact state graphs (utils/contact_state_graph.py:11-16)
  neurals/pretrained_score_function/only_score_model_with_df/2980.pth   LargeScoreFunction weights (neurals/network.py:434-441)
"""
import os
import struct
import sys

import numpy as np

REF = os.environ.get("PREGRASP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "synthesize_pregrasp_b200", "assets")


def read_ply_double(path):
    raw = open(path, "rb").read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode()
    n = int([l for l in header.splitlines() if l.startswith("element vertex")][0].split()[-1])
    props = [l.split()[1:] for l in header.splitlines() if l.startswith("property")]
    assert all(p[0] == "double" for p in props), props
    arr = np.frombuffer(raw[end:end + n * 8 * len(props)], dtype="<f8").reshape(n, len(props))
    return arr.copy()


def read_obj(path):
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(v, np.float64), np.asarray(f, np.int64)


def box_surface_cloud(full_extents, n=2048, seed=0, oversample=6):
    """Deterministic blue-noise cloud on a box surface (stand-in for Open3D's random
    sample_points_poisson_disk, which the reference re-draws at every reset).
    Area-weighted uniform candidates, then farthest-point elimination down to n."""
    rng = np.random.RandomState(seed)
    ex = np.asarray(full_extents, np.float64)
    h = ex / 2
    areas = np.array([ex[1] * ex[2], ex[1] * ex[2], ex[0] * ex[2], ex[0] * ex[2], ex[0] * ex[1], ex[0] * ex[1]])
    m = n * oversample
    face = rng.choice(6, size=m, p=areas / areas.sum())
    uv = rng.uniform(-1, 1, size=(m, 3))
    pts = uv * h
    nrm = np.zeros((m, 3))
    for k in range(6):
        ax, sgn = k // 2, (1.0 if k % 2 == 0 else -1.0)
        sel = face == k
        pts[sel, ax] = sgn * h[ax]
        nrm[sel, ax] = sgn
    keep = [0]
    d = np.linalg.norm(pts - pts[0], axis=1)
    for _ in range(n - 1):
        i = int(np.argmax(d))
        keep.append(i)
        d = np.minimum(d, np.linalg.norm(pts - pts[i], axis=1))
    keep = np.asarray(keep)
    return pts[keep], nrm[keep]


def cylinder_surface_cloud(radius, height, n=2048, seed=0, oversample=6):
    rng = np.random.RandomState(seed)
    a_side, a_cap = 2 * np.pi * radius * height, np.pi * radius ** 2
    m = n * oversample
    which = rng.choice(3, size=m, p=np.array([a_side, a_cap, a_cap]) / (a_side + 2 * a_cap))
    th = rng.uniform(0, 2 * np.pi, m)
    z = rng.uniform(-height / 2, height / 2, m)
    r = radius * np.sqrt(rng.uniform(0, 1, m))
    pts = np.zeros((m, 3)); nrm = np.zeros((m, 3))
    s = which == 0
    pts[s] = np.stack([radius * np.cos(th[s]), radius * np.sin(th[s]), z[s]], 1)
    nrm[s] = np.stack([np.cos(th[s]), np.sin(th[s]), 0 * z[s]], 1)
    for k, sg in ((1, 1.0), (2, -1.0)):
        s = which == k
        pts[s] = np.stack([r[s] * np.cos(th[s]), r[s] * np.sin(th[s]), np.full(s.sum(), sg * height / 2)], 1)
        nrm[s, 2] = sg
    keep = [0]
    d = np.linalg.norm(pts - pts[0], axis=1)
    for _ in range(n - 1):
        i = int(np.argmax(d)); keep.append(i)
        d = np.minimum(d, np.linalg.norm(pts - pts[i], axis=1))
    keep = np.asarray(keep)
    return pts[keep], nrm[keep]


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- score network weights -------------------------------------------------
    import torch
    sd = torch.load(os.path.join(REF, "neurals/pretrained_score_function/only_score_model_with_df/2980.pth"),
                    map_location="cpu")
    w = {k.replace(".", "_"): v.numpy().astype(np.float32) for k, v in sd.items()}
    for k in ("pcn_conv1_weight", "pcn_conv2_weight", "pcn_conv3_weight", "pcn_conv4_weight"):
        w[k] = w[k][:, :, 0]
    np.savez_compressed(os.path.join(OUT, "score_with_df_2980.npz"), **w)
    print({k: v.shape for k, v in w.items()})

    # ---- plate ---------------------------------------------------------------
    ply = read_ply_double(os.path.join(REF, "envs/assets/plate.ply"))
    v, f = read_obj(os.path.join(REF, "envs/assets/plate_cvx_simple.obj"))
    vr = v * 0.9                                   # mesh.scale(0.9, center=0)  plate_env:101
    tn = np.cross(vr[f[:, 1]] - vr[f[:, 0]], vr[f[:, 2]] - vr[f[:, 0]])
    tn /= np.linalg.norm(tn, axis=1, keepdims=True)   # o3d compute_triangle_normals
    np.savez_compressed(os.path.join(OUT, "plate.npz"), cloud=ply[:, :3], cloud_normals=ply[:, 3:6],
                        hull_vertices=v, region_vertices=vr, region_triangles=f, region_normals=tn)

    # ---- box regions -----------------------------------------------------------
    rd = np.load(os.path.join(REF, "data/regions/small_block_dummy_region.npz"))
    np.savez_compressed(os.path.join(OUT, "small_block_region.npz"), regions=rd["regions"],
                        fixed_axes=rd["fixed_axes"], surface_norm=rd["surface_norm"])

    # ---- contact state graphs -----------------------------------------------------
    cs = {}
    root = os.path.join(REF, "data/contact_states")
    for d in sorted(os.listdir(root)):
        env = d.replace("_env", "")
        for fn in ("dummy_states", "csg", "paths_id"):
            p = os.path.join(root, d, fn + ".npy")
            if os.path.exists(p):
                cs[f"{env}__{fn}"] = np.load(p).astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "contact_states.npz"), **cs)

    # ---- deterministic clouds for the primitive-shape objects -------------------------
    clouds = {}
    for env, ext in {"bookshelf": (0.4, 0.4, 0.05), "foodbox": (0.2, 0.4, 0.1), "laptop": (0.24, 0.4, 0.03),
                     "keyboard": (0.24, 0.72, 0.06), "cardboard": (0.24, 0.4, 0.002)}.items():
        p, n = box_surface_cloud(ext)
        clouds[env + "__points"], clouds[env + "__normals"] = p, n
    p, n = cylinder_surface_cloud(0.045, 0.22)      # waterbottle_graph_env.py:176 (mesh r=0.045, physics r=0.041)
    clouds["waterbottle__points"], clouds["waterbottle__normals"] = p, n
    np.savez_compressed(os.path.join(OUT, "primitive_clouds.npz"), **clouds)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
