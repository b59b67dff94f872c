"""Per-environment constant tables ("EnvSpec").

Every literal below is lifted from the reference's env classes and scene builders; the
citation next to each value is the reference file:line it comes from (paths relative
to the reference root).  The reference scatters these through ~740-line gym.Env
classes and PyBullet calls; here they are one flat record that is uploaded once to the
GPU (pg_env_create) and handed, as plain arrays, to the CPU oracle in the tests.

Conventions: quaternions are xyzw (PyBullet order); boxes are (half_extents, centre,
quat); the body creation ORDER is kept because Bullet orders collision pairs by
creation id (it decides which body is "A" in a manifold).
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")

SHAPE_BOX, SHAPE_HULL, SHAPE_CYLINDER = 0, 1, 2
REGION_MESH, REGION_SMALL_BLOCK, REGION_WATERBOTTLE = 0, 1, 2

DT = 1.0 / 250.0            # envs/plate_contact_graph_env.py:86
GRAVITY_Z = -10.0           # envs/plate_contact_graph_env.py:184
CONTROL_SKIP = 50           # envs/plate_contact_graph_env.py:51
NUM_INTERP_F = 7            # model_optimize.py:76
NUM_FINGERS = 2             # model_optimize.py:76 active_finger_tips=[0,1]
SINGLE_ACTION_DIM = NUM_INTERP_F * 3 + 3   # envs/plate_contact_graph_env.py:145
ACTION_DIM = SINGLE_ACTION_DIM * NUM_FINGERS
TIP_MASS = 0.01             # envs/plate_contact_graph_env.py:221
CLOUD_POINTS = 2048         # envs/plate_contact_graph_env.py:203

# envs/scales.py:3-15
SCALES = {
    "bookshelf": (1.0, 1.0, 0.5), "waterbottle": (1.0, 1.0, 1.0), "laptop": (0.6, 1.0, 0.3),
    "foodbox": (0.5, 1.0, 1.0), "plate": (1.0, 1.0, 1.0), "cardboard": (0.6, 1.0, 0.02),
    "keyboard": (1.0, 1.0, 1.0),
}


@dataclasses.dataclass
class StaticBox:
    half: Tuple[float, float, float]
    pos: Tuple[float, float, float]
    quat: Tuple[float, float, float, float]
    friction: float
    name: str = ""


@dataclasses.dataclass
class EnvSpec:
    name: str
    # --- dynamic object -------------------------------------------------------------
    obj_shape: int
    obj_half: Tuple[float, float, float]          # box half extents | (r, r, half_len) for the cylinder
    obj_hull: Optional[np.ndarray]                # [V,3] for SHAPE_HULL (mesh, inside a compound) and SHAPE_CYLINDER (bare prism hull)
    obj_mass: float
    obj_pos: Tuple[float, float, float]
    obj_quat: Tuple[float, float, float, float]
    obj_friction: float
    # --- static scene: floor first, then walls, in creation order -----------------------
    statics: List[StaticBox]
    # creation order of all bodies: "floor", "obj", "wallN"; tips always follow
    order: Sequence[str]
    # --- fingertips -----------------------------------------------------------------
    tip_half_thumb: Tuple[float, float, float]
    tip_half_other: Tuple[float, float, float]
    tip_friction: float
    # --- env logic ------------------------------------------------------------------
    clearance_h: float
    clearance_walls: Sequence[int]                # indices into statics used by checkClearance
    clearance_wall_size: Tuple[float, float, float]
    deocclude_boxes: np.ndarray                   # [B,2,3] (min,max) world AABBs, union (duplicates kept)
    dist_boxes: np.ndarray                        # [M,2,3] (min corner, max corner) of the distance-field meshes
    dist_tris: Optional[np.ndarray]               # [T,3,3] extra triangles (tessellated cylinders) or None
    region_type: int
    settle_substeps: int                          # train-mode substeps after the last control step
    metric_axis: int                              # 2 -> com_h (z), 0 -> com_x
    metric_offset: float
    cloud: np.ndarray                             # [2048,3] object-frame cloud
    cloud_normals: np.ndarray
    # --- region tables ------------------------------------------------------------------
    region_vertices: Optional[np.ndarray] = None  # MeshRegion
    region_triangles: Optional[np.ndarray] = None
    region_normals: Optional[np.ndarray] = None
    block_regions: Optional[np.ndarray] = None    # SmallBlockRegionDummy (already scaled) [30,6]
    block_fixed_axes: Optional[np.ndarray] = None
    block_surface_norm: Optional[np.ndarray] = None
    bottle_radius: float = 0.045                  # utils/waterbottle_region.py:5
    bottle_length: float = 0.22
    # --- contact state graphs -----------------------------------------------------------
    dummy_states: Optional[np.ndarray] = None     # non-validate CSG (model_test.py:86-87)
    csg_states: Optional[np.ndarray] = None       # --validate CSG
    paths_id: Optional[np.ndarray] = None
    dist_prisms: Optional[np.ndarray] = None      # [n,6] (cx, cy, cz, r, half height, sides): dist_tris' polyhedra in closed form

    @property
    def substeps_per_rollout(self) -> int:
        """reset step + N*(pre-simulate + CONTROL_SKIP) + settle, for N=2."""
        return 1 + 2 * (1 + CONTROL_SKIP) + self.settle_substeps


_FLOOR_BIG = dict(half=(150.0, 150.0, 5.0))        # envs/assets/plane.urdf  <box size="300 300 10"/> at z=-5
_FLOOR_SMALL = dict(half=(1.5, 1.5, 5.0))          # envs/assets/small_plane.urdf <box size="3 3 10"/>
_ID = (0.0, 0.0, 0.0, 1.0)
_BIG = 100.0


def _load(name):
    return np.load(os.path.join(_ASSETS, name))


def _box_aabbs(boxes):
    return np.asarray([[b[0], b[1]] for b in boxes], np.float64).reshape(-1, 2, 3)


def _o3d_box(w, d, h, t):
    """o3d.geometry.TriangleMesh.create_box(w,d,h).translate(t): min corner at t."""
    t = np.asarray(t, np.float64)
    return (t, t + np.array([w, d, h], np.float64))


def _o3d_cylinder_tris(radius, height, translate, resolution=20, split=4):
    """Triangles of Open3D's TriangleMesh.create_cylinder(radius,height) default tessellation
    (resolution=20, split=4), z axis, centred, then translated -- the distance field of the
    waterbottle env is measured to this polyhedron (neurals/distancefield_utils.py:62-68)."""
    verts = [np.array([0, 0, height * 0.5]), np.array([0, 0, -height * 0.5])]
    step = 2 * np.pi / resolution
    h_step = height / split
    for i in range(split + 1):
        for j in range(resolution):
            th = step * j
            verts.append(np.array([np.cos(th) * radius, np.sin(th) * radius, height * 0.5 - h_step * i]))
    verts = np.asarray(verts) + np.asarray(translate, np.float64)
    tris = []
    for j in range(resolution):
        j1 = (j + 1) % resolution
        tris.append([0, 2 + j, 2 + j1])
        base = 2 + resolution * split
        tris.append([1, base + j1, base + j])
    for i in range(split):
        b1 = 2 + resolution * i
        b2 = b1 + resolution
        for j in range(resolution):
            j1 = (j + 1) % resolution
            tris.append([b2 + j, b1 + j1, b1 + j])
            tris.append([b2 + j, b2 + j1, b1 + j1])
    return verts[np.asarray(tris)]


def _block_regions(scale):
    rd = _load("small_block_region.npz")
    # utils/small_block_region.py:9  regions * scale.repeat(2)
    return rd["regions"] * np.repeat(np.asarray(scale, np.float64), 2), rd["fixed_axes"], rd["surface_norm"]


def _states(env):
    cs = _load("contact_states.npz")
    get = lambda k: cs[f"{env}__{k}"] if f"{env}__{k}" in cs.files else None
    return get("dummy_states"), get("csg"), get("paths_id")


def _prim_cloud(env):
    c = _load("primitive_clouds.npz")
    return c[env + "__points"], c[env + "__normals"]


def bullet_hull_vertices(obj_vertices: np.ndarray) -> np.ndarray:
    """Vertices of the convex hull PyBullet actually collides with for a URDF mesh (plate_contact_graph_env.py:198).

    The OBJ loader keeps float32 coordinates and the URDF importer runs btConvexHullShape::optimizeConvexHull, whose
    btConvexHullComputer snaps every point to a per-axis integer grid of extent/10216 (truncating toward the centre
    of the bounding box) and returns the snapped coordinates.  The recorded plate trajectory is reproduced to 1e-15
    over the settling phase only with these snapped vertices (3.3e-8 with the raw OBJ numbers).
    """
    v = np.asarray(obj_vertices, np.float64).astype(np.float32).astype(np.float64)
    lo, hi = v.min(0), v.max(0)
    s = (hi - lo) / 10216.0
    c = (lo + hi) * 0.5
    return np.trunc((v - c) * (1.0 / s)) * s + c


def bullet_cylinder_hull(radius: float, half_len: float) -> np.ndarray:
    """What PyBullet collides with for a URDF <cylinder> (waterbottle.urdf): without URDF_USE_IMPLICIT_CYLINDER the
    importer builds a 32-sided prism as a btConvexHullShape -- 64 vertices (r sin a_i, r cos a_i, +-h), a_i = 2 pi i/32
    with i/32 evaluated in float -- and snaps it through optimizeConvexHull like any other hull.  The recorded
    waterbottle trajectory confirms the hull (box-of-AABB inertia), not the analytic cylinder."""
    v = []
    for i in range(32):
        a = 6.283185307179586232 * float(np.float32(i) / np.float32(32))
        v.append([radius * np.sin(a), radius * np.cos(a), half_len])
        v.append([radius * np.sin(a), radius * np.cos(a), -half_len])
    v = np.asarray(v, np.float64)
    lo, hi = v.min(0), v.max(0)
    s = (hi - lo) / 10216.0
    c = (lo + hi) * 0.5
    return np.trunc((v - c) * (1.0 / s)) * s + c


def make_spec(env: str) -> EnvSpec:
    env = {"wallbox": "laptop"}.get(env, env)   # model_optimize.py:25 registers the laptop env as "wallbox"
    floor_dist = _o3d_box(2, 2, 0.02, (-1, -1, -0.02))
    if env == "plate":
        a = _load("plate.npz")
        dummy, csg, paths = _states("plate")
        return EnvSpec(
            name="plate", obj_shape=SHAPE_HULL, obj_half=(0.0, 0.0, 0.0), obj_hull=bullet_hull_vertices(a["hull_vertices"]),
            obj_mass=4.0,                                              # assets/plate_cvx_simple.urdf
            obj_pos=(0.0, 0.0, 0.0), obj_quat=_ID,                     # plate_env:192-199
            obj_friction=100.0,                                        # plate_env:210
            statics=[StaticBox(pos=(0.0, 0.0, -5.0), quat=_ID, friction=70.0, name="floor", **_FLOOR_BIG)],  # :186,:208
            order=("floor", "obj"),
            tip_half_thumb=(0.02, 0.02, 0.011), tip_half_other=(0.01, 0.01, 0.011),   # :218
            tip_friction=150.0,                                        # :227
            clearance_h=0.01, clearance_walls=(), clearance_wall_size=(0, 0, 0),      # :43, :463
            deocclude_boxes=_box_aabbs([((-_BIG, -_BIG, 0.005), (_BIG, _BIG, _BIG))]),  # :662
            dist_boxes=_box_aabbs([floor_dist]), dist_tris=None,        # distancefield_utils.py:70-73
            region_type=REGION_MESH, settle_substeps=100, metric_axis=2, metric_offset=0.0,   # :558-603
            cloud=a["cloud"], cloud_normals=a["cloud_normals"],
            region_vertices=a["region_vertices"], region_triangles=a["region_triangles"],
            region_normals=a["region_normals"], dummy_states=dummy, csg_states=csg, paths_id=paths)

    if env in ("bookshelf", "foodbox", "laptop", "keyboard", "cardboard"):
        sc = SCALES[env]
        pts, nrm = _prim_cloud(env)
        dummy, csg, paths = _states(env)
        common = dict(obj_shape=SHAPE_BOX, obj_hull=None, cloud=pts, cloud_normals=nrm,
                      region_type=REGION_SMALL_BLOCK, dummy_states=dummy, csg_states=csg, paths_id=paths,
                      dist_tris=None)
    if env == "bookshelf":
        br, fa, sn = _block_regions(sc)
        s = 0.7071068
        # envs/setup_pybullet.py:12-34 with scale=(1,1,0.5)
        walls = [
            StaticBox(half=(0.2, 0.2, 0.025), pos=(0.0, 0.055, 0.2), quat=(s, 0, 0, s), friction=0.0, name="w1"),
            StaticBox(half=(0.2, 0.2, 0.025), pos=(0.0, -0.055, 0.2), quat=(s, 0, 0, s), friction=0.0, name="w2"),
            StaticBox(half=(0.05, 0.2, 0.1), pos=(-0.28, 0.0, 0.2), quat=(s, 0, 0, s), friction=0.5, name="w3"),
        ]
        # init orientation: euler XYZ (pi/2, 0, pi/2) -> quaternion (setup_pybullet.py:14-18)
        return EnvSpec(
            name=env, obj_half=(0.2, 0.2, 0.025), obj_mass=1.0,          # assets/book.urdf
            obj_pos=(0.0, 0.0, 0.2), obj_quat=(0.5, -0.5, 0.5, 0.5),   # euler XYZ (pi/2,0,pi/2), setup_pybullet.py:14-18
            obj_friction=70.0,                                           # bookshelf_graph_env.py:204
            statics=[StaticBox(pos=(0.0, 0.0, -5.0), quat=_ID, friction=70.0, name="floor", **_FLOOR_BIG)] + walls,
            order=("floor", "obj", "w1", "w2", "w3"),
            tip_half_thumb=(0.02, 0.02, 0.011), tip_half_other=(0.01, 0.01, 0.011), tip_friction=100.0,  # :213,:222
            clearance_h=0.01, clearance_walls=(1, 2), clearance_wall_size=(0.2, 0.2, 0.05),    # :192, :645-654
            deocclude_boxes=_box_aabbs([((-0.205, -0.045, 0.005), (0.205, 0.045, _BIG)),
                                        ((0.205, -_BIG, 0.005), (_BIG, _BIG, _BIG))]),       # :656-665
            dist_boxes=_box_aabbs([floor_dist,
                                   _o3d_box(0.4, 0.1, 0.4, (-0.2 * sc[0], 0.06 * sc[2], 0.0)),          # right (unscaled!)
                                   _o3d_box(0.4 * sc[0], 0.1 * sc[2], 0.4 * sc[1], (-0.2 * sc[0], -0.16 * sc[2], 0.0))]),
            settle_substeps=300, metric_axis=0, metric_offset=0.0,         # :532-589
            block_regions=br, block_fixed_axes=fa, block_surface_norm=sn, **common)
    if env == "foodbox":
        br, fa, sn = _block_regions(sc)
        return EnvSpec(
            name=env, obj_half=(0.1, 0.2, 0.05), obj_mass=1.0,            # foodbox_env:191
            obj_pos=(0.0, 0.0, 0.05), obj_quat=_ID, obj_friction=70.0,    # :189, :207
            statics=[StaticBox(pos=(0.0, 0.0, -5.0), quat=_ID, friction=70.0, name="floor", **_FLOOR_BIG)],
            order=("floor", "obj"),
            tip_half_thumb=(0.02, 0.02, 0.012), tip_half_other=(0.01, 0.01, 0.01), tip_friction=100.0,  # :214-228
            clearance_h=0.005, clearance_walls=(), clearance_wall_size=(0, 0, 0),
            deocclude_boxes=_box_aabbs([((-_BIG, -_BIG, 0.005), (_BIG, _BIG, _BIG))]),
            dist_boxes=_box_aabbs([floor_dist]),
            settle_substeps=100, metric_axis=2, metric_offset=0.0,
            block_regions=br, block_fixed_axes=fa, block_surface_norm=sn, **common)
    if env == "laptop":
        br, fa, sn = _block_regions(sc)
        wx = -0.11 - 0.2 * sc[0]
        return EnvSpec(
            name=env, obj_half=(0.12, 0.2, 0.015), obj_mass=1.0,          # laptop_env:195
            obj_pos=(0.0, 0.0, 0.05 * sc[2] * sc[2]), obj_quat=_ID,       # :193,:197 (SCALE applied twice)
            obj_friction=10.0,                                            # :212
            statics=[StaticBox(pos=(0.0, 0.0, -5.0), quat=_ID, friction=0.1, name="floor", **_FLOOR_BIG),
                     StaticBox(half=(0.1, 2.0, 0.5), pos=(wx, 0.0, 0.5), quat=_ID, friction=0.0, name="w1")],  # :184-187
            order=("floor", "w1", "obj"),
            tip_half_thumb=(0.01, 0.02, 0.011), tip_half_other=(0.01, 0.01, 0.01), tip_friction=100.0,  # :222,:231
            clearance_h=0.005, clearance_walls=(), clearance_wall_size=(0, 0, 0),
            deocclude_boxes=_box_aabbs([((wx, -_BIG, 0.005), (_BIG, _BIG, _BIG))]),       # :664
            dist_boxes=_box_aabbs([floor_dist, _o3d_box(2, 0.02, 1, (-1, -0.22, 0))]),     # distancefield_utils.py:29-34
            settle_substeps=100, metric_axis=2, metric_offset=0.0,
            block_regions=br, block_fixed_axes=fa, block_surface_norm=sn, **common)
    if env == "keyboard":
        br, fa, sn = _block_regions((0.6, 1.8, 0.6))                      # keyboard_env:103
        return EnvSpec(
            name=env, obj_half=(0.12, 0.36, 0.03), obj_mass=3.0,          # assets/keyboard.urdf
            obj_pos=(0.38, 0.0, 0.03), obj_quat=_ID, obj_friction=10.0,   # setup_pybullet.py:59, keyboard_env:202
            statics=[StaticBox(pos=(-1.0, 0.0, -5.0), quat=_ID, friction=0.1, name="floor", **_FLOOR_SMALL),  # :58,:199
                     StaticBox(half=(0.1, 2.0, 0.5), pos=(-0.31, 0.0, 0.5), quat=_ID, friction=0.0, name="w1")],  # :62-65
            order=("floor", "obj", "w1"),
            tip_half_thumb=(0.01, 0.01, 0.01), tip_half_other=(0.01, 0.01, 0.01), tip_friction=5.0,   # :211,:220
            clearance_h=0.005, clearance_walls=(), clearance_wall_size=(0, 0, 0),
            deocclude_boxes=_box_aabbs([((-0.31, -_BIG, 0.005), (0.5, _BIG, _BIG)),
                                        ((0.5, -_BIG, -_BIG), (_BIG, _BIG, _BIG))]),      # envs/bounding_boxes.py:18-19
            dist_boxes=_box_aabbs([_o3d_box(3, 3, 0.02, (-2.5, -1.5, -0.02)),
                                   _o3d_box(0.02, 2, 1, (-0.22, -1, 0))]),               # distancefield_utils.py:36-41
            settle_substeps=100, metric_axis=0, metric_offset=-0.38,        # keyboard_env:550,594
            block_regions=br, block_fixed_axes=fa, block_surface_norm=sn, **common)
    if env == "cardboard":
        br, fa, sn = _block_regions(sc)
        s = 0.7071068
        return EnvSpec(
            name=env, obj_half=(0.12, 0.2, 0.001), obj_mass=2.0,          # assets/cardboard.urdf
            obj_pos=(0.0, 0.2 * sc[0], 0.001), obj_quat=(0.0, 0.0, -s, s),  # setup_pybullet.py:139-140
            obj_friction=70.0,                                            # cardboard_env:196
            statics=[StaticBox(pos=(0.0, 1.48, -5.0), quat=_ID, friction=0.002, name="floor", **_FLOOR_SMALL)],  # :138,:194
            order=("floor", "obj"),
            tip_half_thumb=(0.02, 0.02, 0.012), tip_half_other=(0.01, 0.01, 0.012), tip_friction=100.0,  # :203,:212
            clearance_h=0.001, clearance_walls=(), clearance_wall_size=(0, 0, 0),
            deocclude_boxes=_box_aabbs([((-1.5, -0.02, 0.0005), (1.5, 3.0, _BIG)),
                                        ((-1.5, -_BIG, -_BIG), (1.5, -0.02, _BIG))]),     # :644-657
            dist_boxes=_box_aabbs([_o3d_box(3, 3, 10, (-1.5, -0.02, -10))]),               # distancefield_utils.py:94-97
            settle_substeps=100, metric_axis=2, metric_offset=0.0,
            block_regions=br, block_fixed_axes=fa, block_surface_norm=sn, **common)
    if env == "waterbottle":
        pts, nrm = _prim_cloud(env)
        dummy, csg, paths = _states(env)
        s = 0.7071068
        walls = [  # envs/setup_pybullet.py:85-98
            StaticBox(half=(0.04, 0.025, 0.12), pos=(0.0, -0.075, 0.12), quat=_ID, friction=0.0, name="w1"),
            StaticBox(half=(0.04, 0.025, 0.12), pos=(0.0, 0.075, 0.12), quat=_ID, friction=0.0, name="w2"),
            StaticBox(half=(0.05, 5.0, 10.0), pos=(-0.1, 0.0, 5.0), quat=(s, 0, 0, s), friction=0.5, name="w3"),
        ]
        tris = np.concatenate([_o3d_cylinder_tris(0.045, 0.22, (0, -0.1, 0.11)),
                               _o3d_cylinder_tris(0.045, 0.22, (0, 0.1, 0.11))])
        return EnvSpec(
            name=env, obj_shape=SHAPE_CYLINDER, obj_half=(0.041, 0.041, 0.11),
            obj_hull=bullet_cylinder_hull(0.041, 0.11),                                       # assets/waterbottle.urdf
            obj_mass=1.0, obj_pos=(0.0, 0.0, 0.115), obj_quat=_ID, obj_friction=70.0,        # setup_pybullet.py:87
            statics=[StaticBox(pos=(0.0, 0.0, -5.0), quat=_ID, friction=70.0, name="floor", **_FLOOR_BIG)] + walls,
            order=("floor", "obj", "w1", "w2", "w3"),
            tip_half_thumb=(0.01, 0.01, 0.011), tip_half_other=(0.01, 0.01, 0.011), tip_friction=100.0,  # :204,:213
            clearance_h=0.01, clearance_walls=(1, 2), clearance_wall_size=(0.045, 0.045, 0.11),  # :191,:616-625
            deocclude_boxes=_box_aabbs([((-0.04, -0.045, 0.005), (0.05, 0.045, _BIG)),
                                        ((0.05, -_BIG, 0.005), (_BIG, _BIG, _BIG))]),        # :633-636
            dist_boxes=_box_aabbs([floor_dist, _o3d_box(0.1, 0.4, 0.4, (-0.15, -0.2, 0))]), dist_tris=tris,
            # the two tessellated holders in closed form (20-sided prisms, Open3D's default resolution): what the kernels evaluate;
            # the triangle list above is what the oracle evaluates
            dist_prisms=np.array([[0, -0.1, 0.11, 0.045, 0.11, 20], [0, 0.1, 0.11, 0.045, 0.11, 20]], np.float64),
            region_type=REGION_WATERBOTTLE, settle_substeps=300, metric_axis=0, metric_offset=0.0,   # :507-560
            cloud=pts, cloud_normals=nrm, dummy_states=dummy, csg_states=csg, paths_id=paths)
    raise KeyError(f"unknown env {env!r}")


ENV_NAMES = ("plate", "bookshelf", "foodbox", "laptop", "keyboard", "cardboard", "waterbottle")
