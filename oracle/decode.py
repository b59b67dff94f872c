"""Action decode of the contact-graph envs, numpy fp64 (test oracle, see oracle/__init__.py).

Follows, line by line, envs/plate_contact_graph_env.py:269-391 (step prologue,
transform_a_to_action_single_pc, calc_force_interp_from_fvec, transform_a_no_opt_time),
the region parsers utils/mesh_region.py:34-51, utils/small_block_region.py:18-48,
utils/waterbottle_region.py:9-40, utils/contact_state_graph.py:37-38 and
utils/math_utils.py:457-474 (rotation_matrix_from_vectors).
"""
import numpy as np

DT = 1.0 / 250.0
CONTROL_SKIP = 50
NUM_INTERP = 7
SINGLE = NUM_INTERP * 3 + 3


def rotation_matrix_from_vectors(vec1, vec2):
    # utils/math_utils.py:457-474
    a = (vec1 / np.linalg.norm(vec1)).reshape(3)
    b = (vec2 / np.linalg.norm(vec2)).reshape(3)
    v = np.cross(a, b)
    c = np.dot(a, b)
    s = np.linalg.norm(v)
    if s == 0:
        if c > 0:
            return np.eye(3)
        return np.array([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]])
    k = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + k + k.dot(k) * ((1 - c) / (s ** 2))


def parse_mesh_region(region_id, sub_action, vertices, triangles, normals):
    # utils/mesh_region.py:34-51, mode "uniform" (:19-23)
    sa = (sub_action + 1) * 0.5
    vtx = vertices[triangles[region_id]]
    sx = np.sqrt(sa[0])
    w = np.array([1 - sx, sx * (1 - sa[1]), sx * sa[1]])
    res = np.zeros(3)
    res += vtx[0] * w[0]
    res += vtx[1] * w[1]
    res += vtx[2] * w[2]
    return res, normals[region_id]


def parse_small_block_region(region_id, sub_action, regions, fixed_axes, surface_norm):
    # utils/small_block_region.py:18-48
    sa = (sub_action + 1) * 0.5
    fa = fixed_axes[region_id]
    r = regions[region_id]
    if fa == 0:
        x = r[0]
        y = (r[3] - r[2]) * sa[0] + r[2]
        z = (r[5] - r[4]) * sa[1] + r[4]
    elif fa == 1:
        x = (r[1] - r[0]) * sa[0] + r[0]
        y = r[2]
        z = (r[5] - r[4]) * sa[1] + r[4]
    else:
        x = (r[1] - r[0]) * sa[0] + r[0]
        y = (r[3] - r[2]) * sa[1] + r[2]
        z = r[4]
    return np.array([x, y, z]), surface_norm[region_id]


def parse_waterbottle_region(region_id, sub_action, radius=0.045, length=0.22):
    # utils/waterbottle_region.py:9-40
    sa = (sub_action + 1) * 0.5
    if region_id == 0 or region_id == 3:
        r = radius * np.sqrt(sa[0])
        th = sa[1] * 2 * np.pi
        x, y = -r * np.sin(th), -r * np.cos(th)
        z = length / 2 if region_id == 0 else -length / 2
        n = np.array([0.0, 0.0, 1.0 if region_id == 0 else -1.0])
    elif region_id == 1 or region_id == 2:
        z = sa[0] * length - length / 2
        th = sa[1] * np.pi + (np.pi if region_id == 2 else 0.0)
        x, y = -radius * np.sin(th), -radius * np.cos(th)
        n = np.array([x, y, 0])
        n = n / np.linalg.norm(n)
    else:
        raise IndexError(region_id)
    return np.array([x, y, z]), n


class Regions:
    """Plain-array view of one env's contact-region tables (built by the tests from the
    product's EnvSpec fields -- the oracle itself imports nothing from the product)."""

    def __init__(self, region_type, csg_states, **tables):
        self.region_type = int(region_type)     # 0 mesh, 1 small block, 2 waterbottle
        self.csg = np.asarray(csg_states)
        self.t = tables

    def parse(self, state_id, finger, sub_action):
        rid = int(self.csg[state_id][finger]) - 1          # csg.getState(state)[finger] - 1
        if self.region_type == 0:
            n = len(self.t["triangles"])
            return parse_mesh_region(rid % n if rid < 0 else rid, sub_action, self.t["vertices"],
                                     self.t["triangles"], self.t["normals"])
        if self.region_type == 1:
            n = len(self.t["regions"])
            return parse_small_block_region(rid % n if rid < 0 else rid, sub_action, self.t["regions"],
                                            self.t["fixed_axes"], self.t["surface_norm"])
        return parse_waterbottle_region(rid, sub_action)


def new_carry(num_fingers=2):
    """last_fins (pos, on) and last_surface_norm after reset (plate_env:233, :245-249)."""
    return {"pos": np.zeros((num_fingers, 3)), "on": np.zeros(num_fingers, bool),
            "norm": np.zeros((num_fingers, 3))}


def decode_step(action, carry, regions, path, step_cnt, num_fingers=2):
    """One env.step() prologue.  Returns (cur_pos[F,3], cur_on[F], f_interp[F,3,50]) and
    updates `carry` in place the way step() does at its end (last_fins = cur_fins)."""
    a = np.tanh(np.asarray(action, np.float64))                       # plate_env:487
    state = int(path[step_cnt])
    prev = None if step_cnt == 0 else int(path[step_cnt - 1])       # :269-270
    cur_pos = np.zeros((num_fingers, 3))
    cur_on = np.zeros(num_fingers, bool)
    f_interp = np.zeros((num_fingers, 3, CONTROL_SKIP))
    x_interp = np.arange(CONTROL_SKIP, dtype=np.float64)
    x_d = np.linspace(0, float(CONTROL_SKIP), NUM_INTERP, endpoint=True)
    new_norm = carry["norm"].copy()
    for f in range(num_fingers):
        sub = a[f * SINGLE:(f + 1) * SINGLE]
        if carry["on"][f]:                                           # :290-305
            pos = carry["pos"][f].copy()
            nrm = carry["norm"][f]
        else:                                                        # :307-310
            pos, nrm = regions.parse(state, f, sub[-3:-1])
            pos = pos + 0.01 * nrm
            new_norm[f] = nrm
        R = rotation_matrix_from_vectors(np.array([0.0, 0.0, 1.0]), nrm)
        knots = np.zeros((NUM_INTERP, 3))
        for i in range(NUM_INTERP):                                  # :313-330
            if sub[3 * i] > 0:
                vn = sub[3 * i] * 5.0 * DT
                vt1 = sub[3 * i + 1] * 1.0 * vn
                vt2 = sub[3 * i + 2] * 1.0 * vn
            else:
                vn = vt1 = vt2 = 0.0
            knots[i] = -vn * nrm + R @ np.array([vt1, vt2, 0])
        broke = (prev is not None and regions.csg[state][f] != regions.csg[prev][f]) and carry["on"][f]
        flag = 0.0 if broke else (1.0 if sub[-1] > 0 else 0.0)       # :336-339
        if flag == 0:                                                # :380-382
            knots[:] = 0.0
            pos = np.array([100.0, 100.0, 100.0])
        if knots[0, 0] == 0:                                         # :384-385
            knots[:] = 0.0
        for row in range(3):                                         # :346-362
            f_interp[f, row] = np.interp(x_interp, x_d, knots[:, row])
        cur_pos[f] = pos
        cur_on[f] = flag > 0
    carry["pos"], carry["on"], carry["norm"] = cur_pos.copy(), cur_on.copy(), new_norm
    return cur_pos, cur_on, f_interp
