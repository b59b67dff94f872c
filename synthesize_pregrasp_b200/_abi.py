"""ctypes view of include/pregrasp_b200.h and the loader of the in-tree CUDA library.

The library is the product: if it is missing or fails to load there is NO fallback -- every entry point
raises.  `make_c_spec` turns an EnvSpec (envspec.py) into the PgEnvSpec struct pg_env_create takes.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import envspec as ES

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpregrasp_b200.so")

PG_MAX_STATICS = 4


class PgBox(C.Structure):
    _fields_ = [("half", C.c_double * 3), ("pos", C.c_double * 3), ("quat", C.c_double * 4), ("friction", C.c_double),
                ("compound", C.c_int32), ("pad", C.c_int32)]


_FP = C.POINTER(C.c_float)
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int32)

_W_NAMES = ["conv1_w", "conv1_b", "conv2_w", "conv2_b", "conv3_w", "conv3_b", "conv4_w", "conv4_b",
            "fc1_w", "fc1_b", "fc2_w", "fc2_b", "fc3_w", "fc3_b", "out_w", "out_b"]


class PgScoreWeights(C.Structure):
    _fields_ = [(n, _FP) for n in _W_NAMES]


class PgEnvSpec(C.Structure):
    _fields_ = [
        ("obj_shape", C.c_int32), ("n_hull", C.c_int32), ("obj_half", C.c_double * 3), ("hull", _DP),
        ("obj_mass", C.c_double), ("obj_pos", C.c_double * 3), ("obj_quat", C.c_double * 4), ("obj_friction", C.c_double),
        ("n_static", C.c_int32), ("floor_index", C.c_int32), ("statics", PgBox * PG_MAX_STATICS),
        ("tip_half", (C.c_double * 3) * 2), ("tip_friction", C.c_double),
        ("clearance_h", C.c_double), ("clearance_wall_size", C.c_double * 3),
        ("n_clear_walls", C.c_int32), ("clear_walls", C.c_int32 * 2),
        ("project_uses_clearance", C.c_int32), ("settle_substeps", C.c_int32), ("metric_axis", C.c_int32),
        ("metric_offset", C.c_double),
        ("region_type", C.c_int32), ("n_regions", C.c_int32), ("n_states", C.c_int32), ("pad0", C.c_int32),
        ("tri_v", _DP), ("region_n", _DP), ("block_r", _DP), ("block_axis", _IP), ("csg", _IP),
        ("n_cloud", C.c_int32), ("n_crop", C.c_int32), ("cloud", _DP),
        ("crop", ((C.c_double * 3) * 2) * 2),
        ("n_dist_box", C.c_int32), ("n_dist_tri", C.c_int32),
        ("dist_box", ((C.c_double * 3) * 2) * 4), ("dist_tri", _DP),
        ("n_dist_prism", C.c_int32), ("pad1", C.c_int32), ("dist_prism", (C.c_double * 6) * 2),
        ("weights", PgScoreWeights),
    ]


def load_score_weights(path=None):
    path = path or os.path.join(_HERE, "assets", "score_with_df_2980.npz")
    w = np.load(path)
    ren = lambda k: k.replace("pcn_", "").replace("_weight", "_w").replace("_bias", "_b")
    return {ren(k): np.ascontiguousarray(w[k], np.float32) for k in w.files}


def make_c_spec(spec: ES.EnvSpec, states=None, weights=None):
    """-> (PgEnvSpec, keepalive).  `states` selects the contact state graph ([S,2] region ids, 1-based);
    default = the env's dummy_states (model_test.py:86-87); pass spec.csg_states for --validate runs."""
    keep = []

    def dp(a):
        a = np.ascontiguousarray(a, np.float64)
        keep.append(a)
        return a.ctypes.data_as(_DP)

    def ip(a):
        a = np.ascontiguousarray(a, np.int32)
        keep.append(a)
        return a.ctypes.data_as(_IP)

    s = PgEnvSpec()
    s.obj_shape = spec.obj_shape
    s.obj_half[:] = spec.obj_half
    if spec.obj_hull is not None:
        s.n_hull = len(spec.obj_hull)
        s.hull = dp(spec.obj_hull)
    s.obj_mass = spec.obj_mass
    s.obj_pos[:] = spec.obj_pos
    s.obj_quat[:] = spec.obj_quat
    s.obj_friction = spec.obj_friction
    # statics in CREATION order among themselves (spec.order lists floor/walls/obj)
    names = [n for n in spec.order if n != "obj"]
    by_name = {b.name: b for b in spec.statics}
    s.n_static = len(names)
    for i, n in enumerate(names):
        b = by_name[n]
        s.statics[i].half[:] = b.half
        s.statics[i].pos[:] = b.pos
        s.statics[i].quat[:] = b.quat
        s.statics[i].friction = b.friction
        s.statics[i].compound = 1 if n == "floor" else 0
    s.floor_index = names.index("floor")
    s.tip_half[0][:] = spec.tip_half_thumb
    s.tip_half[1][:] = spec.tip_half_other
    s.tip_friction = spec.tip_friction
    s.clearance_h = spec.clearance_h
    s.clearance_wall_size[:] = spec.clearance_wall_size
    s.n_clear_walls = len(spec.clearance_walls)
    for i, wi in enumerate(spec.clearance_walls):
        s.clear_walls[i] = names.index(spec.statics[wi].name)
    s.project_uses_clearance = 1 if len(spec.clearance_walls) else 0
    s.settle_substeps = spec.settle_substeps
    s.metric_axis = spec.metric_axis
    s.metric_offset = spec.metric_offset
    states = spec.dummy_states if states is None else np.asarray(states)
    s.region_type = spec.region_type
    s.n_states = len(states)
    s.csg = ip(states[:, :2])
    if spec.region_type == ES.REGION_MESH:
        s.n_regions = len(spec.region_triangles)
        s.tri_v = dp(spec.region_vertices[spec.region_triangles])
        s.region_n = dp(spec.region_normals)
    elif spec.region_type == ES.REGION_SMALL_BLOCK:
        s.n_regions = len(spec.block_regions)
        s.block_r = dp(spec.block_regions)
        s.block_axis = ip(spec.block_fixed_axes)
        s.region_n = dp(spec.block_surface_norm)
    else:
        s.n_regions = 4
    s.n_cloud = len(spec.cloud)
    s.cloud = dp(spec.cloud)
    s.n_crop = len(spec.deocclude_boxes)
    for b in range(s.n_crop):
        for k in range(2):
            s.crop[b][k][:] = spec.deocclude_boxes[b][k]
    s.n_dist_box = len(spec.dist_boxes)
    for b in range(s.n_dist_box):
        for k in range(2):
            s.dist_box[b][k][:] = spec.dist_boxes[b][k]
    if getattr(spec, "dist_prisms", None) is not None:          # closed form of the same polyhedra: the triangle list stays with the oracle
        s.n_dist_prism = len(spec.dist_prisms)
        for i, pr in enumerate(spec.dist_prisms):
            for j in range(6):
                s.dist_prism[i][j] = float(pr[j])
    elif spec.dist_tris is not None:
        s.n_dist_tri = len(spec.dist_tris)
        s.dist_tri = dp(spec.dist_tris)
    weights = weights or load_score_weights()
    for n in _W_NAMES:
        a = np.ascontiguousarray(weights[n], np.float32)
        keep.append(a)
        setattr(s.weights, n, a.ctypes.data_as(_FP))
    return s, keep


_lib = None


def lib():
    """Load libpregrasp_b200.so (built by __graft_entry__.build / csrc/Makefile).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PREGRASP_B200_LIB", LIB_PATH)     # override: kernel-variant experiments only
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, f64 = C.c_void_p, C.c_int, C.c_double
    L.pg_last_error.restype = C.c_char_p
    L.pg_version.restype = C.c_char_p
    L.pg_launch_count.restype = C.c_uint64
    L.pg_env_create.argtypes = [C.POINTER(PgEnvSpec), i32, C.POINTER(vp)]
    L.pg_env_destroy.argtypes = [vp]
    L.pg_rollout.argtypes = [vp, vp, vp, i32, i32, _IP, f64, vp, vp, vp, vp]
    L.pg_rollout_multi.argtypes = [vp, vp, vp, i32, i32, i32, vp, _IP, f64, vp, vp, vp]
    L.pg_rollout_host.argtypes = [vp, _DP, _FP, i32, i32, _IP, f64, _DP, _DP, _DP]
    L.pg_physics.argtypes = [vp, vp, vp, i32, i32, _IP, f64, vp, vp, vp, vp, vp, vp]
    L.pg_cost.argtypes = [vp, vp, vp, i32, vp, vp]
    L.pg_score.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.pg_mppi_update.argtypes = [vp, vp, i32, i32, i32, f64, f64, vp, vp, i32, vp]
    L.pg_mppi_merge_apply.argtypes = [_DP, i32, i32, i32, f64, f64, _DP]
    L.pg_dyn_feasible_simple.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.pg_dyn_sweep_simple.argtypes = [vp, vp, i32, f64, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp]
    L.pg_dyn_feasible_qp.argtypes = [vp, vp, vp, i32, i32, f64, f64, f64, vp, vp, vp]
    L.pg_dyn_sweep_qp.argtypes = [vp, vp, i32, f64, f64, f64, f64, C.c_uint64, C.c_uint64, vp, C.c_uint64, vp, vp]
    L.pg_set_rollout_mode.argtypes = [vp, i32]
    L.pg_carry_create.argtypes = [vp, C.POINTER(vp)]
    L.pg_carry_reset.argtypes = [vp, _DP]
    L.pg_carry_destroy.argtypes = [vp]
    L.pg_env_step_single.argtypes = [vp, vp, _DP, i32, f64, i32, _DP, _DP, _DP, _DP]
    for f in ("pg_env_create", "pg_env_destroy", "pg_rollout", "pg_rollout_multi", "pg_rollout_host", "pg_physics", "pg_cost", "pg_score",
              "pg_mppi_update", "pg_mppi_merge_apply", "pg_dyn_feasible_simple", "pg_dyn_sweep_simple", "pg_dyn_feasible_qp",
              "pg_dyn_sweep_qp"):
        getattr(L, f).restype = C.c_int
    _lib = L
    return L


EXPORTS = ("pg_last_error", "pg_version", "pg_env_create", "pg_env_destroy", "pg_rollout", "pg_rollout_multi", "pg_rollout_host", "pg_physics",
           "pg_cost", "pg_score", "pg_mppi_update", "pg_mppi_merge_apply", "pg_dyn_feasible_simple", "pg_dyn_sweep_simple",
           "pg_dyn_feasible_qp", "pg_dyn_sweep_qp", "pg_carry_create", "pg_carry_reset", "pg_carry_destroy", "pg_env_step_single", "pg_set_rollout_mode",
           "pg_launch_count")


def check(rc):
    if rc != 0:
        raise RuntimeError(f"pregrasp_b200 error {rc}: {lib().pg_last_error().decode()}")
