"""Host model of warp-level solver cost: how much would re-sorting rollouts between phases save?
Runs the product's per-thread rollout logic on the host (tests/hostsim) with random noise and records the solver
iteration count of every substep; warp cost = sum over substeps of the max over its 32 lanes."""
import ctypes as C, os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from synthesize_pregrasp_b200 import _abi
from synthesize_pregrasp_b200.envspec import make_spec
from concurrent.futures import ProcessPoolExecutor

ENV, F, K = sys.argv[1] if len(sys.argv) > 1 else "bookshelf", 50.0, int(sys.argv[2]) if len(sys.argv) > 2 else 1024

def run(seed_k):
    seed, k = seed_k
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    H = C.CDLL(so)
    spec = make_spec(ENV)
    cs, keep = _abi.make_c_spec(spec)
    rng = np.random.RandomState(seed)
    noise = (0.8 * rng.randn(k, 2, 48)).astype(np.float32)
    u = np.zeros((2, 48))
    N = 2
    pose = np.zeros((k, N, 7)); fins = np.zeros((k, N, 2, 3)); metric = np.zeros(k)
    iters = np.zeros((k, N * 50), np.int32); settle = np.zeros((k, 512), np.int32)
    path = np.zeros(N, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    H.hostsim_set_settle_buffer(vp(settle))
    assert H.hostsim_rollout(C.byref(cs), vp(u), vp(noise), k, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric), None, vp(iters)) == 0
    return iters, settle

if __name__ == "__main__":
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio", "-o", so,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    with ProcessPoolExecutor(8) as ex:
        parts = list(ex.map(run, [(s, K // 8) for s in range(8)]))
    iters = np.concatenate([p[0] for p in parts]); settle = np.concatenate([p[1] for p in parts])
    ns = int((settle.sum(0) > 0).sum())
    settle = settle[:, :ns]
    print("rollouts", len(iters), "control substeps", iters.shape[1], "settle substeps", ns)
    print("mean iterations: control %.1f settle %.1f" % (iters.mean(), settle.mean()))
    def warp_cost(mat, order):
        m = mat[order]
        m = m[: len(m) // 32 * 32].reshape(-1, 32, mat.shape[1])
        return m.max(axis=1).sum()
    ident = np.arange(len(iters))
    base = warp_cost(iters, ident) + warp_cost(settle, ident)
    lane = (iters.sum() + settle.sum()) / 32.0
    print("baseline warp cost %.0f; lane-perfect %.0f (%.2fx)" % (base, lane, base / lane))
    c1, c2 = iters[:, :50], iters[:, 50:]
    for name, key in (("last control substep", c2[:, -1]), ("mean of last 10 control substeps", c2[:, -10:].mean(1)),
                      ("mean of step 2", c2.mean(1)), ("true settle cost (bound)", settle.sum(1))):
        o = np.argsort(key, kind="stable")
        t = warp_cost(iters, ident) + warp_cost(settle, o)
        print("settle sorted by %-36s %.0f (%.3fx)" % (name, t, base / t))
    o1 = np.argsort(c1.mean(1), kind="stable")
    t = warp_cost(c1, ident) + warp_cost(c2, o1) + warp_cost(settle, np.argsort(c2[:, -10:].mean(1), kind="stable"))
    print("step 2 sorted by step-1 mean, settle by last-10 mean: %.0f (%.3fx)" % (t, base / t))
    print("share of warp cost in settle: %.2f" % (warp_cost(settle, ident) / base))
    # CTA-local variants: groups of G rollouts sorted among themselves
    def local_sort(key, G):
        o = np.arange(len(key))
        for g in range(0, len(key), G):
            o[g:g + G] = g + np.argsort(key[g:g + G], kind="stable")
        return o
    for G in (448, 256, 128):
        t = warp_cost(iters, ident) + warp_cost(settle, local_sort(c2.mean(1), G))
        print("settle sorted inside groups of %d by step-2 mean: %.3fx" % (G, base / t))
    # periodic re-sorts during the settle phase (key = mean of the previous `period` substeps)
    for G in (448, 1 << 30):
        for period in (25, 50, 100):
            t = warp_cost(iters, ident)
            key = c2.mean(1)
            for s0 in range(0, ns, period):
                seg = settle[:, s0:s0 + period]
                t += warp_cost(seg, local_sort(key, min(G, len(key))))
                key = seg[:, -min(period, 25):].mean(1)
            print("group %s, re-sort every %d settle substeps: %.3fx" % ("all" if G > 1e6 else G, period, base / t))
    # control phase too: re-sort at the start of step 2 by step-1 mean and every `period` substeps
    for period in (25, 50):
        allm = np.concatenate([iters, settle], axis=1)
        t = warp_cost(allm[:, :period], ident)
        key = allm[:, :period].mean(1)
        for s0 in range(period, allm.shape[1], period):
            seg = allm[:, s0:s0 + period]
            t += warp_cost(seg, local_sort(key, 448))
            key = seg.mean(1)
        print("group 448, re-sort every %d substeps from the start: %.3fx" % (period, base / t))
