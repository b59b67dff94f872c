"""Can the order in which PyBullet's re-created object proxy finds its pairs be DERIVED?  An experiment (CPU, seconds per recording).

A plain-Python btDbvt (insertleaf / removeleaf / update / collideTV with Bullet's stack order, optimizeIncremental WITHOUT the
pointer-comparing `sort`) is driven by the oracle's replay of a recording: every body's AABB is mirrored into the tree exactly where
the env's calls would touch Bullet's broadphase (addCollisionObject at reset, setCollisionFilterPair = destroyProxy + createProxy of
both bodies, updateAabbs of every step with gDbvtMargin = 0: a leaf whose AABB changed is removed and re-inserted at the root).  At the
LAST re-creation of the object's proxy in each env step the script prints the order in which collideTV meets the floor and the
fingertips -- to be compared with the order each recording needs (env step 1 floor first everywhere; env step 2
thumb first for cardboard2/3, plate_ablation_df, plate_no_df, floor first for keyboard_scale, cardboard_scale, keyboard, cardboard,
plate).
  python tools/whatif/dbvt_find_order.py [recording ...]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PHYS, oracle_spec, recorded, recording_time_spec   # noqa: E402
from oracle import envsim                                               # noqa: E402


class Node:
    __slots__ = ("lo", "hi", "parent", "ch", "data")

    def __init__(self, lo, hi, parent=None, data=None):
        self.lo, self.hi, self.parent, self.ch, self.data = np.array(lo, float), np.array(hi, float), parent, None, data

    def leaf(self):
        return self.ch is None


def _merge(a, b):
    return np.minimum(a.lo, b.lo), np.maximum(a.hi, b.hi)


def _prox(a, b):                                       # Proximity(): Manhattan distance of the doubled centres
    return np.abs((a.lo + a.hi) - (b.lo + b.hi)).sum()


def _contain(a, lo, hi):
    return bool(np.all(a.lo <= lo) and np.all(a.hi >= hi))


def _intersect(n, lo, hi):
    return bool(np.all(n.lo <= hi) and np.all(n.hi >= lo))


class Dbvt:
    def __init__(self):
        self.root, self.opath, self.leaves = None, 0, 0

    def insertleaf(self, root, leaf):
        if self.root is None:
            self.root, leaf.parent = leaf, None
            return
        while not root.leaf():
            root = root.ch[0 if _prox(leaf, root.ch[0]) < _prox(leaf, root.ch[1]) else 1]     # Select()
        prev = root.parent
        lo, hi = _merge(leaf, root)
        node = Node(lo, hi, prev)
        node.ch = [root, leaf]
        root.parent = leaf.parent = node
        if prev is not None:
            prev.ch[1 if prev.ch[1] is root else 0] = node
            while prev is not None:
                if not _contain(prev, node.lo, node.hi):
                    prev.lo, prev.hi = _merge(prev.ch[0], prev.ch[1])
                else:
                    break
                node, prev = prev, prev.parent
        else:
            self.root = node

    def removeleaf(self, leaf):
        if leaf is self.root:
            self.root = None
            return None
        parent = leaf.parent
        prev = parent.parent
        sib = parent.ch[0 if parent.ch[1] is leaf else 1]
        if prev is not None:
            prev.ch[1 if prev.ch[1] is parent else 0] = sib
            sib.parent = prev
            while prev is not None:
                lo, hi = prev.lo.copy(), prev.hi.copy()
                prev.lo, prev.hi = _merge(prev.ch[0], prev.ch[1])
                if np.any(lo != prev.lo) or np.any(hi != prev.hi):
                    prev = prev.parent
                else:
                    break
            return prev if prev is not None else self.root
        self.root, sib.parent = sib, None
        return self.root

    def insert(self, lo, hi, data):
        leaf = Node(lo, hi, None, data)
        self.insertleaf(self.root, leaf)
        self.leaves += 1
        return leaf

    def remove(self, leaf):
        self.removeleaf(leaf)
        self.leaves -= 1

    def update(self, leaf, lo=None, hi=None):
        root = self.removeleaf(leaf)
        if root is not None:
            root = self.root                               # lookahead -1
        if lo is not None:
            leaf.lo, leaf.hi = np.array(lo, float), np.array(hi, float)
        self.insertleaf(root, leaf)

    def optimize_incremental(self, passes=1):
        if self.root is None:
            return
        for _ in range(passes):
            node, bit = self.root, 0
            while not node.leaf():
                node = node.ch[(self.opath >> bit) & 1]    # (Bullet: sort(node, root) first -- compares node ADDRESSES; left out)
                bit = (bit + 1) & 31
            self.update(node)
            self.opath += 1

    def collide_tv(self, lo, hi, skip=None):
        out = []
        if self.root is None:
            return out
        stack = [self.root]
        while stack:
            n = stack.pop()
            if _intersect(n, lo, hi):
                if n.leaf():
                    if n is not skip:
                        out.append(n.data)
                else:
                    stack.append(n.ch[0]); stack.append(n.ch[1])
        return out


RESET_UPDATES_AABB = int(os.environ.get("RESET_UPDATES_AABB", "0"))
REVERSE_OBJECTS = int(os.environ.get("REVERSE_OBJECTS", "0"))


class Mirror:
    """ctypes-library proxy: forwards every call to the oracle and mirrors the broadphase side effects into a Dbvt"""

    def __init__(self, L, sim, margin02=0.02):
        self.L, self.sim, self.t, self.leaf, self.m = L, sim, Dbvt(), {}, margin02
        self.found = []                                 # (world step of the NEXT pgo_step, body, [found bodies]) per createProxy
        self.nstep = 0

    def __getattr__(self, k):
        return getattr(self.L, k)

    def _aabb(self, h, body, extra):
        out = np.zeros(6)
        self.L.pgo_get_aabb(h, body, ctypes.c_double(extra), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        return out[:3], out[3:]

    def pgo_add_body(self, h, *a):
        b = self.L.pgo_add_body(h, *a)
        lo, hi = self._aabb(h, b, 0.0)                   # addCollisionObject: createProxy with the shape's own AABB
        self.leaf[b] = self.t.insert(lo, hi, b)
        return b

    def _recreate(self, h, b):
        self.t.remove(self.leaf[b])
        lo, hi = self._aabb(h, b, 0.0)
        self.leaf[b] = self.t.insert(lo, hi, b)
        self.found.append((self.nstep + 1, b, self.t.collide_tv(lo, hi, self.leaf[b])))

    def pgo_set_filter(self, h, a, b, enable):
        self.L.pgo_set_filter(h, a, b, enable)
        if self.nstep > 0:                               # (the filter calls of reset() happen before the proxies matter)
            self._recreate(h, a); self._recreate(h, b)

    def pgo_reset_pose(self, h, b, pos, quat):
        self.L.pgo_reset_pose(h, b, pos, quat)
        if RESET_UPDATES_AABB:                           # resetBasePositionAndOrientation -> updateSingleAabb (teleport)
            lo, hi = self._aabb(h, b, self.m)
            if not _contain(self.leaf[b], lo, hi):
                self.t.update(self.leaf[b], lo, hi)

    def pgo_step(self, h):
        for b in sorted(self.leaf, reverse=bool(REVERSE_OBJECTS)):   # updateAabbs, collision-object order; gDbvtMargin = 0
            lo, hi = self._aabb(h, b, self.m)
            lf = self.leaf[b]
            if not _contain(lf, lo, hi):
                self.t.update(lf, lo, hi)
        self.t.optimize_incremental(1)                   # btDbvtBroadphase::collide
        self.L.pgo_step(h)
        self.nstep += 1


def run(exp):
    g = recorded(exp)
    osd = oracle_spec(recording_time_spec(exp))
    sim = envsim.EnvSim(osd, g["max_force"])
    L = envsim.lib()
    L.pgo_get_aabb.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
    mir = Mirror(L, sim)
    sim.L = mir
    sim.reset()
    for j in range(2):
        sim.step(g["u"][j], [0, 0, 0])
    names = {sim.obj: "obj", sim.floor: "floor", sim.tips[0]: "thumb", sim.tips[1]: "index"}
    for k, t in enumerate(sim.idle):
        names[t] = f"idle{k}"
    last = {}
    for step, body, found in mir.found:
        if body == sim.obj:
            last[step] = found
    out = []
    for step in sorted(last):
        if step in (3, 54):                              # the last re-creation before the first servo substep of env step 1 / 2
            out.append((step, [names.get(b, f"wall{b}") for b in last[step]]))
    return out


if __name__ == "__main__":
    for exp in (sys.argv[1:] or list(PHYS)):
        print(f"{exp:26s}", "   ".join(f"world step {s}: " + " < ".join(f) for s, f in run(exp)))
