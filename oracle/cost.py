"""Per-step cost inputs (test oracle, see oracle/__init__.py).

Restates envs/plate_contact_graph_env.py:393-437 (calc_reward_value_stage_one), :661-684
(deocclude / project_to_pcd / compute_distance_field), the multi-box variants
envs/bookshelf_graph_env.py:645-690, neurals/distancefield_utils.py:15-22 and
utils/math_utils.py:476-490 (minimumDist2dBox).  Open3D is replaced by its definitions:
crop = closed AABB test, KD-tree 1-NN = exact nearest point, RaycastingScene.compute_distance =
exact unsigned point-to-mesh distance evaluated on float32 query points.
"""
import numpy as np


def quat_to_matrix_xyzw(q):
    # utils/helper.py:304-307 -> pytorch_kinematics.quaternion_to_matrix on (w,x,y,z)
    x, y, z, w = [float(v) for v in q]
    two_s = 2.0 / (w * w + x * x + y * y + z * z)
    return np.array([
        [1 - two_s * (y * y + z * z), two_s * (x * y - z * w), two_s * (x * z + y * w)],
        [two_s * (x * y + z * w), 1 - two_s * (x * x + z * z), two_s * (y * z - x * w)],
        [two_s * (x * z - y * w), two_s * (y * z + x * w), 1 - two_s * (x * x + y * y)]])


def box_unsigned_distance(p, lo, hi):
    """Exact unsigned distance from points p[N,3] to the SURFACE of the solid box [lo,hi]
    (= distance to its 12-triangle mesh): outside -> distance to the box, inside -> to the nearest face."""
    d_out = np.maximum(np.maximum(lo - p, p - hi), 0.0)
    outside = np.linalg.norm(d_out, axis=1)
    inside = np.minimum(p - lo, hi - p).min(axis=1)
    return np.where(outside > 0, outside, np.maximum(inside, 0.0))


def point_triangle_distance(p, tri):
    """Exact distance from points p[N,3] to triangles tri[T,3,3] -> [N] (min over T).  Ericson 5.1.5."""
    best = np.full(len(p), np.inf)
    for a, b, c in tri:
        ab, ac, ap = b - a, c - a, p - a
        d1, d2 = ap @ ab, ap @ ac
        bp = p - b
        d3, d4 = bp @ ab, bp @ ac
        cp = p - c
        d5, d6 = cp @ ab, cp @ ac
        vc = d1 * d4 - d3 * d2
        vb = d5 * d2 - d1 * d6
        va = d3 * d6 - d5 * d4
        q = np.empty_like(p)
        done = np.zeros(len(p), bool)

        def put(mask, val):
            m = mask & ~done
            q[m] = val[m] if val.ndim == 2 else val
            done[m] = True
        put((d1 <= 0) & (d2 <= 0), a)
        put((d3 >= 0) & (d4 <= d3), b)
        with np.errstate(divide="ignore", invalid="ignore"):
            put((vc <= 0) & (d1 >= 0) & (d3 <= 0), a + (d1 / (d1 - d3))[:, None] * ab)
            put((d6 >= 0) & (d5 <= d6), c)
            put((vb <= 0) & (d2 >= 0) & (d6 <= 0), a + (d2 / (d2 - d6))[:, None] * ac)
            put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0),
                b + ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[:, None] * (c - b))
            den = 1.0 / (va + vb + vc)
            put(np.ones(len(p), bool), a + ab * (vb * den)[:, None] + ac * (vc * den)[:, None])
        best = np.minimum(best, np.linalg.norm(p - q, axis=1))
    return best


def minimum_dist_2d_box(point, box_pos, box_quat_xyzw, shape):
    # utils/math_utils.py:476-490
    R = quat_to_matrix_xyzw(box_quat_xyzw)
    x, y, z = R.T @ (np.asarray(point, np.float64) - np.asarray(box_pos, np.float64))
    d = [max(abs(x) - shape[0], 0), max(abs(y) - shape[1], 0), max(abs(z) - shape[2], 0)]
    return float(np.linalg.norm(d))


def check_clearance(point, clearance_h, walls, wall_size):
    """walls: list of (pos, quat) of the static boxes the env tests (bookshelf_graph_env.py:645-654).
    With no walls this is the plate rule `z < CLEARANCE_H -> not clear` (plate_env:463)."""
    if point[2] < clearance_h:
        return False
    for pos, quat in walls:
        if minimum_dist_2d_box(point, pos, quat, wall_size) < clearance_h:
            return False
    return True


def project_to_pcd(local_point, cloud, clearance_h, walls, wall_size, strict_z):
    """plate_env:668-674 (`point[2] > CLEARANCE_H`, strict_z=True) and bookshelf_graph_env.py:683-689
    (`checkClearance(point)` applied to the LOCAL point -- reproduced as written)."""
    ok = (local_point[2] > clearance_h) if strict_z else check_clearance(local_point, clearance_h, walls, wall_size)
    if not ok:
        return np.array([100.0, 100.0, 100.0])
    d = np.linalg.norm(cloud - np.asarray(local_point), axis=1)
    return cloud[int(np.argmin(d))].copy()


def cost_inputs(pos, quat, cloud, deocclude_boxes, dist_boxes, dist_tris=None):
    """-> (pts_obj float32 [P',3], df float32 [P'], keep index list).  Points inside several crop
    boxes appear once per box, in box order, as the reference's vstack does."""
    R = quat_to_matrix_xyzw(quat)
    world = cloud @ R.T + np.asarray(pos, np.float64)                  # rotate(R, center=0); translate(pos)
    idx = []
    for lo, hi in deocclude_boxes:
        inside = np.all((world >= lo) & (world <= hi), axis=1)
        idx.append(np.nonzero(inside)[0])
    idx = np.concatenate(idx) if idx else np.zeros(0, int)
    w = world[idx]
    q32 = w.astype(np.float32).astype(np.float64)                      # compute_distance(points.astype(float32))
    df = np.full(len(w), np.inf)
    for lo, hi in dist_boxes:
        df = np.minimum(df, box_unsigned_distance(q32, lo, hi))
    if dist_tris is not None and len(w):
        df = np.minimum(df, point_triangle_distance(q32, dist_tris))
    back = (w - np.asarray(pos, np.float64)) @ R                       # translate(-pos); rotate(R.T)
    return back.astype(np.float32), df.astype(np.float32), idx
