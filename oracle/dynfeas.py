"""Dynamic-feasibility checks (test oracle, see oracle/__init__.py).

simple_check_*: restates utils/dyn_feasibility_check.py:124-235 (project2norm, check_in_triangle incl. its
`c1<0 and c1<0` typo, check_sign_change, simple_check_dyn_feasible_3p, simple_check_dyn_feasible); the two
pytorch_kinematics calls at :191-196 (axis_angle_to_quaternion, quaternion_apply) are restated from their
published definitions (pytorch3d transforms) in float64.
qp_matrices: the QP data of :28-60 / :93-101 (Q, G, h and the 6x9 wrench map).  qp_objective / check_dyn_feasible_qp solve that
QP -- the reference's own (Q, G, h) formulation -- with scipy's SLSQP instead of cvxpy/OSQP.  PARITY UNPINNED: nothing in the
reference records a QP optimum and its solver cannot be installed, so this pins the kernel to an independent exact solve of the
same problem, not to the reference's solver output (which is only accurate to OSQP's tolerance anyway).
"""
import itertools

import numpy as np

MU = 5.0                  # DEFAULT_FRICTION_COEFF :20
MIN_NORMAL_FORCE_3 = 1.0  # MIN_NORMAL_FORCE[3] :15-17
MIN_DISTANCE_BETWEEN_FINGER = 0.03  # :19


def _axis_angle_to_quaternion(aa):
    ang = np.linalg.norm(aa)
    half = 0.5 * ang
    if abs(ang) < 1e-6:
        s = 0.5 - (ang * ang) / 48
    else:
        s = np.sin(half) / ang
    return np.array([np.cos(half), aa[0] * s, aa[1] * s, aa[2] * s])


def _qmul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def _quaternion_apply(q, p):
    pq = np.array([0.0, p[0], p[1], p[2]])
    return _qmul(_qmul(q, pq), q * np.array([1.0, -1, -1, -1]))[1:]


def project2norm(v, n):
    return v.dot(n) / (np.linalg.norm(n) ** 2) * n


def check_in_triangle(vertices, point):
    res0 = np.cross(vertices[1] - vertices[0], point - vertices[0])
    res1 = np.cross(vertices[2] - vertices[1], point - vertices[1])
    res2 = np.cross(vertices[0] - vertices[2], point - vertices[2])
    c1 = res0.dot(res1)
    c2 = res0.dot(res2)
    return bool((c1 > 0 and c2 > 0) or (c1 < 0 and c1 < 0))


def check_sign_change(vectors, vertices):
    center = (vertices[0] + vertices[1] + vertices[2]) / 3
    prev_c = None
    for i in range(len(vertices)):
        r = vertices[i] - center
        c1 = np.cross(r, vectors[2 * i])
        c2 = np.cross(r, vectors[2 * i + 1])
        if prev_c is None:
            prev_c = c1
        if prev_c.dot(c1) < 0 or c1.dot(c2) < 0:
            return True
        prev_c = c2
    return False


def simple_check_dyn_feasible_3p(pts, nrm):
    with np.errstate(all="ignore"):
        v1 = pts[1] - pts[0]
        v2 = pts[2] - pts[0]
        n = np.cross(v1, v2)
        if np.linalg.norm(n) < 1e-4:
            return False
        n = n / np.linalg.norm(n)
        pj_n = [project2norm(nrm[i], n) for i in range(3)]
        pj_p = [nrm[i] - pj_n[i] for i in range(3)]
        for i in range(3):
            if np.linalg.norm(pj_n[i]) / np.linalg.norm(pj_p[i]) > MU:
                return False
        cones = []
        for i in range(3):
            temp = np.linalg.norm(nrm[i]) / np.linalg.norm(pj_p[i]) * np.linalg.norm(pj_n[i])
            angle = np.arctan(np.sqrt(MU ** 2 - temp ** 2) / np.sqrt(temp ** 2 + np.linalg.norm(nrm[i]) ** 2))
            p = pj_p[i] / np.linalg.norm(pj_p[i])
            cones.append(_quaternion_apply(_axis_angle_to_quaternion(n * angle), p))
            cones.append(_quaternion_apply(_axis_angle_to_quaternion(-n * angle), p))
        c00, c01, c10, c11, c20, c21 = cones
        origin = np.zeros(3)
        ok = False
        for tri in ([c00, c10, c20], [c00, c11, c20], [c00, c10, c21], [c00, c11, c21],
                    [c01, c10, c20], [c01, c11, c20], [c01, c10, c21], [c01, c11, c21]):
            if check_in_triangle(tri, origin):
                ok = True
                break
        if not ok:
            return False
        return bool(check_sign_change(cones, pts))


def simple_check_dyn_feasible(pts, nrm):
    pts, nrm = np.asarray(pts, np.float64), np.asarray(nrm, np.float64)
    if len(pts) == 3:
        return simple_check_dyn_feasible_3p(pts, nrm)
    for c in itertools.combinations(range(len(pts)), 3):
        c = list(c)
        if simple_check_dyn_feasible_3p(pts[c], nrm[c]):
            return True
    return False


# ------------------------------------------------------------------------------------------------------
def qp_matrices(p, nrm_unit_or_raw, mu=MU, normalise=True):
    """Q, G, h of get_min_external_wrench_qp_matrices (:28-60) for one triple; contact frames as in
    check_dyn_feasible (:93-101, n normalised) or find_dynamically_feasible_contacts (:277-285, raw n)."""
    C = np.zeros((3, 3, 3))
    for i in range(3):
        n = nrm_unit_or_raw[i] / np.linalg.norm(nrm_unit_or_raw[i]) if normalise else nrm_unit_or_raw[i]
        t0 = np.array([n[1], -n[0], 0.0])
        C[i, :, 0], C[i, :, 1], C[i, :, 2] = n, t0, np.cross(n, t0)
    skew = lambda v: np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    A1 = np.vstack([C[i].T for i in range(3)])                 # 9x3
    A2 = np.vstack([(skew(p[i]) @ C[i]).T for i in range(3)])  # 9x3
    Q = A1 @ A1.T + A2 @ A2.T
    G = np.zeros((15, 9))
    for i in range(3):
        G[4 * i:4 * i + 4, 3 * i:3 * i + 3] = [[mu, -1, 0], [mu, 0, -1], [mu, 1, 0], [mu, 0, 1]]
        G[12 + i, 3 * i] = 1.0
    h = np.hstack([np.zeros(12), -np.ones(3)]) * MIN_NORMAL_FORCE_3
    W = np.hstack([A1, A2]).T                                   # 6x9 wrench map, Q = W^T W
    return Q, G, h, W


def qp_objective(p, nrm, normalise=True, mu=MU):
    """Optimal value of construct_and_solve_wrench_closure_qp (:62-75) for one triple, by SLSQP on (Q, G, h)."""
    from scipy.optimize import minimize
    Q, G, h, _ = qp_matrices(np.asarray(p, float), np.asarray(nrm, float), mu, normalise)
    z0 = np.tile([-2.0 * MIN_NORMAL_FORCE_3, 0.0, 0.0], 3)
    r = minimize(lambda z: 0.5 * z @ Q @ z, z0, jac=lambda z: Q @ z, method="SLSQP", options={"ftol": 1e-15, "maxiter": 1000},
                 constraints=[{"type": "ineq", "fun": lambda z: h - G @ z, "jac": lambda z: -G}])
    return float(r.fun)


def check_dyn_feasible_qp(points, normals, tol=2e-3):
    """check_dyn_feasible (:77-112): first triple of the present contacts (x < 50) whose QP optimum is <= tol.
    -> (decision, smallest objective met before deciding)."""
    import itertools
    points, normals = np.asarray(points, float), np.asarray(normals, float)
    keep = points[:, 0] < 50
    points, normals = points[keep], normals[keep]
    best = np.inf
    for comb in itertools.combinations(range(len(points)), 3):
        o = qp_objective(points[list(comb)], normals[list(comb)], normalise=True)
        best = min(best, o)
        if o <= tol:
            return True, best
    return False, best
