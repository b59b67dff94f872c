"""Single-rollout env simulation on top of the C++ physics oracle (test infrastructure).

Mirrors the control flow of envs/plate_contact_graph_env.py: reset() :176-250, step() :253-632
(pre_simulate :454-482, servo loop :504-533, settle :556-603) with every `self._p.*` call mapped
onto oracle/physics/bullet_restatement.cpp.  The spec arrives as a plain dict of arrays/values
(built by tests from the product's EnvSpec) -- the oracle imports nothing from the product.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import decode as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "physics", "libpgo_physics.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "physics")], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        L.pgo_create.restype = ctypes.c_void_p
        dp = ctypes.POINTER(ctypes.c_double)
        L.pgo_destroy.argtypes = [ctypes.c_void_p]
        L.pgo_set_param.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.pgo_add_body.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, ctypes.c_int, ctypes.c_double, dp, dp,
                                   ctypes.c_double, ctypes.c_int, ctypes.c_double]
        L.pgo_set_roles.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3
        L.pgo_add_idle.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.pgo_set_filter.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 3
        L.pgo_reset_pose.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp]
        L.pgo_get_pose.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        L.pgo_get_vel.argtypes = [ctypes.c_void_p, ctypes.c_int, dp]
        L.pgo_create_constraint.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dp, dp]
        L.pgo_change_constraint.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, ctypes.c_double]
        L.pgo_remove_constraint.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.pgo_b3_multiply_transforms.argtypes = [dp, dp, dp, dp]
        L.pgo_b3_multiply_transforms.restype = None
        L.pgo_step.argtypes = [ctypes.c_void_p]
        L.pgo_last_iterations.argtypes = [ctypes.c_void_p]
        L.pgo_dump_manifolds.argtypes = [ctypes.c_void_p, dp, ctypes.c_int]
        _LIB = L
    return _LIB


def _p(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def qmat(q):
    """btMatrix3x3::setRotation(xyzw)."""
    x, y, z, w = q
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy], [xy + wz, 1.0 - (xx + zz), yz - wx],
                     [xz - wy, yz + wx, 1.0 - (xx + yy)]])


def b3_multiply_transforms_native(lib, posA, ornA, posB):
    """The same arithmetic through the C++ oracle library (oracle/physics pgo_b3_multiply_transforms): used inside EnvSim so that
    the float32 emulation does not slow the CPU baseline down (the numpy version costs 40 % of a rollout)."""
    a = np.ascontiguousarray(posA, np.float64); q = np.ascontiguousarray(ornA, np.float64); b = np.ascontiguousarray(posB, np.float64)
    out = np.empty(3)
    P = ctypes.POINTER(ctypes.c_double)
    lib.pgo_b3_multiply_transforms(a.ctypes.data_as(P), q.ctypes.data_as(P), b.ctypes.data_as(P), out.ctypes.data_as(P))
    return out


def b3_multiply_transforms(posA, ornA, posB):
    """pybullet.multiplyTransforms(posA, ornA, posB, [0,0,0,1])[0] as PyBullet computes it: b3Transform with b3Scalar = FLOAT
    (PhysicsClientC_API.cpp b3MultiplyTransforms; the extension is built with double-precision Bullet but single-precision
    Bullet3Common).  Pinned by data/tip_data/*: this reproduces the recorded fingertip positions bit for bit (every row of the
    foodbox, keyboard, laptop, waterbottle and cardboard recordings), the double-precision product only to 1e-8.  The env feeds
    these float32 results into the physics: teleport position / constraint pivot (plate_env:462-474) and servo targets (:526-530)."""
    f = np.float32
    x, y, z, w = (f(v) for v in ornA)
    d = f(f(f(f(x * x) + f(y * y)) + f(z * z)) + f(w * w))            # b3Quaternion::length2 = dot(q, q)
    s = f(f(2.0) / d)                                                 # b3Matrix3x3::setRotation
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
    for i in range(3):                                                # b3Transform::operator(): dot3 (left to right) + origin
        acc = f(f(f(v[0] * R[i][0]) + f(v[1] * R[i][1])) + f(v[2] * R[i][2]))
        out[i] = f(acc + o[i])
    return out


_ZERO3 = np.zeros(3)


class EnvSim:
    """spec keys: obj_shape, obj_half, obj_hull, obj_mass, obj_pos, obj_quat, obj_friction,
    statics=[(half,pos,quat,friction)], order=[names], static_names, tip_half=[thumb, other], tip_friction,
    clearance_h, clearance_walls, clearance_wall_size, settle_substeps, regions (oracle.decode.Regions)"""

    def __init__(self, spec, max_force, params=None):
        self.s = spec
        self.max_force = float(max_force)
        self.params = dict(params or {})
        self.f32_api = bool(self.params.pop("f32_api", 1))         # pybullet.multiplyTransforms in float32 (see b3_multiply_transforms)
        self.h = None
        self.L = lib()

    def close(self):
        if self.h is not None:
            self.L.pgo_destroy(self.h)
            self.h = None

    __del__ = close

    # ------------------------------------------------------------------
    def _add(self, shape, half, verts, mass, pos, quat, friction, compound=False, margin=0.001):
        a1, hp = _p(half)
        v = np.zeros((0, 3)) if verts is None else np.asarray(verts, np.float64)
        a2, vp = _p(v)
        a3, pp = _p(pos)
        a4, qp = _p(quat)
        return self.L.pgo_add_body(self.h, int(shape), hp, vp, len(v), float(mass), pp, qp, float(friction),
                                   int(compound), float(margin))

    def pose(self, body):
        out = np.zeros(7)
        self.L.pgo_get_pose(self.h, body, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out

    def vel(self, body):
        out = np.zeros(6)
        self.L.pgo_get_vel(self.h, body, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out

    def manifolds(self):
        buf = np.zeros(35 * 32)
        n = self.L.pgo_dump_manifolds(self.h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(buf))
        out = []
        for k in range(0, n, 35):
            r = buf[k:k + 35]
            out.append((int(r[0]), int(r[1]), int(r[2]), r[3:].reshape(4, 8)[:int(r[2])].copy()))
        return out

    def reset(self):
        s = self.s
        self.close()
        self.h = self.L.pgo_create()
        for k, v in self.params.items():
            if k in ("obj_compound", "idle_tips"):
                continue
            self.L.pgo_set_param(self.h, k.encode(), float(v))
        self.static_ids = {}
        names = s["static_names"]
        for name in s["order"]:
            if name == "obj":
                # shape 1: URDF mesh = hull inside a btCompoundShape; shape 2: URDF <cylinder> = bare 64-vertex prism hull
                shape = 1 if (s["obj_shape"] == 2 and s.get("obj_hull") is not None) else s["obj_shape"]
                self.obj = self._add(shape, s["obj_half"], s.get("obj_hull"), s["obj_mass"], s["obj_pos"],
                                     s["obj_quat"], s["obj_friction"],
                                     compound=bool(self.params.get("obj_compound", s["obj_shape"] == 1)))
            else:
                half, pos, quat, fr = s["statics"][names.index(name)]
                self.static_ids[name] = self._add(0, half, None, 0.0, pos, quat, fr, compound=(name == "floor"))
        self.floor = self.static_ids["floor"]
        self.tips = []
        for f in range(2):                                   # plate_env:215-228 (idle tips 2-4 never interact)
            half = s["tip_half"][0] if f == 0 else s["tip_half"][1]
            t = self._add(0, half, None, 0.01, (f + 2.0,) * 3, (0, 0, 0, 1), s["tip_friction"])
            self.tips.append(t)
            self.L.pgo_set_filter(self.h, t, self.floor, 0)
        # the env's inactive fingertips (plate_env:215-228: `all_fingers` = 0..4, only `active_finger_tips` = [0, 1] are ever
        # moved): fingers 2-4 are created at the SAME place (100, 100, 100), so they collide with each other, are pushed apart and
        # fall freely.  They never touch anything else, but their manifolds sit in the dispatcher's array and take part in the
        # unstable quickSort by island id that fixes the solver's manifold order -- until they have drifted apart (world steps 20
        # and 39), after which the order of everybody else's manifolds flips (keyboard_20.0 row 37).
        self.idle = []
        for f in range(2, 2 + int(self.params.get("idle_tips", 3))):
            t = self._add(0, s["tip_half"][1], None, 0.01, (100.0, 100.0, 100.0), (0, 0, 0, 1), s["tip_friction"])
            self.idle.append(t)
            self.L.pgo_set_filter(self.h, t, self.floor, 0)
            self.L.pgo_add_idle(self.h, t)
        self.L.pgo_set_roles(self.h, self.obj, self.tips[0], self.tips[1])
        self.carry = D.new_carry()
        self.L.pgo_step(self.h)                              # :237
        self.step_cnt = 0
        self.cons_active = False
        return self.pose(self.obj)

    def _clear(self, p):
        s = self.s
        if p[2] < s["clearance_h"]:
            return False
        from .cost import minimum_dist_2d_box
        for wi in s["clearance_walls"]:
            half, pos, quat, fr = s["statics"][wi]
            if minimum_dist_2d_box(p, pos, quat, s["clearance_wall_size"]) < s["clearance_h"]:
                return False
        return True

    def step(self, action, path, record=True, train=False, n_steps=2):
        """-> dict(poses[50,7] at the START of each control substep, cur_pos, cur_on, metric)"""
        s, L, h = self.s, self.L, self.h
        cur_pos, cur_on, f_interp = D.decode_step(action, self.carry, s["regions"], path, self.step_cnt)
        self.step_cnt += 1
        o = self.pose(self.obj)
        pos, quat = o[:3], o[3:]
        R = qmat(quat)
        # ---- pre_simulate :454-482
        if self.cons_active:
            for i in range(2):
                L.pgo_remove_constraint(h, i)
        for i in range(2):
            g = b3_multiply_transforms_native(L, pos, quat, cur_pos[i]) if self.f32_api else pos + R @ cur_pos[i]
            if not self._clear(g):
                g = np.array([100.0, 100.0, 100.0])
            L.pgo_set_filter(h, self.tips[i], self.obj, 0)
            a1, gp = _p(g)
            a2, qp = _p(quat)
            L.pgo_reset_pose(h, self.tips[i], gp, qp)
            L.pgo_create_constraint(h, i, self.tips[i], gp, qp)
        self.cons_active = True
        L.pgo_step(h)
        for i in range(2):
            L.pgo_set_filter(h, self.tips[i], self.obj, 1)
        # ---- servo loop :504-533
        poses = np.zeros((50, 7))
        iters = np.zeros(50, int)
        for t in range(50):
            o = self.pose(self.obj)
            poses[t] = o
            pos, quat = o[:3], o[3:]
            R = qmat(quat)
            vel = self.vel(self.obj)[:3]
            if self.f32_api:                                   # both fingers in one native call (pgo_servo_targets)
                DP = ctypes.POINTER(ctypes.c_double)
                f6 = np.ascontiguousarray(np.stack([f_interp[0][:, t], f_interp[1][:, t]]))
                c6 = np.ascontiguousarray(cur_pos, np.float64)
                v3 = np.ascontiguousarray(vel)
                o7 = np.ascontiguousarray(o)
                tars = np.empty((2, 3))
                L.pgo_servo_targets(o7.ctypes.data_as(DP), c6.ctypes.data_as(DP), f6.ctypes.data_as(DP), v3.ctypes.data_as(DP),
                                    ctypes.c_double(float(D.DT)), tars.ctypes.data_as(DP))
            for i in range(2):
                if self.f32_api:
                    tar = tars[i]
                else:
                    v_g = R @ f_interp[i][:, t] + vel * D.DT
                    tar = (pos + R @ cur_pos[i]) + v_g
                a1, tp = _p(tar)
                a2, qp = _p(quat)
                L.pgo_change_constraint(h, i, tp, qp, self.max_force)
            L.pgo_step(h)
            iters[t] = L.pgo_last_iterations(h)
        metric = 0.0
        if self.step_cnt == n_steps and train:
            ax = s.get("metric_axis", 2)
            metric = (self.pose(self.obj)[ax] + s.get("metric_offset", 0.0)) * 0.5
            for _ in range(s["settle_substeps"]):
                L.pgo_step(h)
            metric += (self.pose(self.obj)[ax] + s.get("metric_offset", 0.0)) * 0.5
        return dict(poses=poses, cur_pos=cur_pos, cur_on=cur_on, metric=metric, iters=iters,
                    final_pose=self.pose(self.obj))
