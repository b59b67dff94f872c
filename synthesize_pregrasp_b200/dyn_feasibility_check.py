"""GPU entry points with the names and argument meaning of utils/dyn_feasibility_check.py:
simple_check_dyn_feasible (:226-235) and check_dyn_feasible (:77-112, the wrench-closure QP) for batches, and the exhaustive
triple sweep of find_dynamically_feasible_contacts (:239-303) with either criterion."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from .engine import _ptr, _require_cuda, _stream

MIN_DISTANCE_BETWEEN_FINGER = 0.03   # utils/dyn_feasibility_check.py:19
DEFAULT_FRICTION_COEFF = 5           # :20
MIN_NORMAL_FORCE = {3: 1.0, 4: 1000.0}   # :14-16 (the QP is always built for a triple: n_fingers = 3)


def simple_check_dyn_feasible(contact_points, contact_normals, count=None, device=None):
    """contact_points / contact_normals: [M,3] (one grasp -> bool, like the reference) or [B,M,3] (-> bool array).
    `count` [B] gives the number of valid contacts per item when they are padded to a common M."""
    _require_cuda()
    L = _abi.lib()
    p = np.asarray(contact_points, np.float64)
    n = np.asarray(contact_normals, np.float64)
    single = p.ndim == 2
    if single:
        p, n = p[None], n[None]
    B, M = p.shape[0], p.shape[1]
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    pt, nt = torch.as_tensor(p, device=dev).contiguous(), torch.as_tensor(n, device=dev).contiguous()
    ct = torch.as_tensor(np.asarray(count, np.int32), device=dev).contiguous() if count is not None else None
    out = torch.empty(B, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _abi.check(L.pg_dyn_feasible_simple(_ptr(pt), _ptr(nt), _ptr(ct), B, M, _ptr(out), _stream()))
    res = out.cpu().numpy().astype(bool)
    return bool(res[0]) if single else res


def check_dyn_feasible(contact_points, contact_normals, tol=2e-3, count=None, device=None, return_objective=False):
    """The QP criterion (:77-112): some contact triple admits friction-pyramid forces (mu = 5, normal force >= 1) whose net
    wrench has 1/2 |w|^2 <= tol.  [M,3] (one grasp) or [B,M,3]; contacts with x >= 50 are the reference's "absent finger"
    marker (:80) and are dropped.  -> bool / bool array (+ the smallest objective met when return_objective).
    The reference returns the contact rows of the first passing triple or None; this returns the decision."""
    _require_cuda()
    L = _abi.lib()
    p = np.asarray(contact_points, np.float64)
    n = np.asarray(contact_normals, np.float64)
    single = p.ndim == 2
    if single:
        p, n = p[None], n[None]
    B, M = p.shape[0], p.shape[1]
    if count is None:                                            # compact the present contacts to the front (:80-81)
        keep = p[:, :, 0] < 50
        order = np.argsort(~keep, axis=1, kind="stable")
        p = np.take_along_axis(p, order[:, :, None], 1)
        n = np.take_along_axis(n, order[:, :, None], 1)
        count = keep.sum(1)
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    pt, nt = torch.as_tensor(p, device=dev).contiguous(), torch.as_tensor(n, device=dev).contiguous()
    ct = torch.as_tensor(np.asarray(count, np.int32), device=dev).contiguous()
    out = torch.empty(B, dtype=torch.uint8, device=dev)
    obj = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _abi.check(L.pg_dyn_feasible_qp(_ptr(pt), _ptr(nt), _ptr(ct), B, M, float(DEFAULT_FRICTION_COEFF), float(MIN_NORMAL_FORCE[3]),
                                        float(tol), _ptr(out), _ptr(obj), _stream()))
    res = out.cpu().numpy().astype(bool)
    o = obj.cpu().numpy()
    if single:
        return (bool(res[0]), float(o[0])) if return_objective else bool(res[0])
    return (res, o) if return_objective else res


def triple_range(rank, world_size, n_points):
    """[first, last) of the lexicographic triple index space C(n_points, 3) owned by `rank` of `world_size` GPUs: contiguous,
    disjoint, covering, sizes differing by at most one (no collective except the final count / gather)."""
    total = n_points * (n_points - 1) * (n_points - 2) // 6
    base, rem = divmod(total, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def find_dynamically_feasible_contacts(points, normals, min_dist=MIN_DISTANCE_BETWEEN_FINGER, first=0, last=None,
                                       max_out=1 << 22, device=None, criterion="simple", tol=1e-4):
    """All index triples (i<j<k) of the cloud passing the pairwise-distance prefilter (:260-265) and the closure test;
    criterion "simple" = the geometric test of simple_check_dyn_feasible, "qp" = the wrench-closure QP with `tol`, which is
    what the reference function itself runs (:274-279).  [first,last) selects a slice of the lexicographic triple index
    space (one per GPU).  -> (triples [n,3] int32 in unspecified order, total count)."""
    _require_cuda()
    L = _abi.lib()
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    pt = torch.as_tensor(np.asarray(points, np.float64), device=dev).contiguous()
    nt = torch.as_tensor(np.asarray(normals, np.float64), device=dev).contiguous()
    P = pt.shape[0]
    total = P * (P - 1) * (P - 2) // 6
    if last is None and first == 0:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():         # one slice of the index space per rank
            first, last = triple_range(dist.get_rank(), dist.get_world_size(), P)
    last = total if last is None else min(last, total)
    out = torch.empty((max_out, 3), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        if criterion == "qp":
            _abi.check(L.pg_dyn_sweep_qp(_ptr(pt), _ptr(nt), P, float(min_dist), float(DEFAULT_FRICTION_COEFF), float(MIN_NORMAL_FORCE[3]),
                                         float(tol), C.c_uint64(first), C.c_uint64(last), _ptr(out), C.c_uint64(max_out), _ptr(cnt), _stream()))
        elif criterion == "simple":
            _abi.check(L.pg_dyn_sweep_simple(_ptr(pt), _ptr(nt), P, float(min_dist), C.c_uint64(first), C.c_uint64(last),
                                             _ptr(out), C.c_uint64(max_out), _ptr(cnt), _stream()))
        else:
            raise ValueError(f"unknown criterion {criterion!r}")
    n = int(cnt.item())
    return out[:min(n, max_out)].cpu().numpy(), n
