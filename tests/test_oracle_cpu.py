"""CPU tests: the oracle against every golden vector / recorded fixture the reference offers for the path,
the host logic, and that the C-ABI library loads and exports every symbol include/pregrasp_b200.h declares."""
import ctypes as C
import os
import sys
import re
import subprocess

import numpy as np
import pytest

from helpers import HORIZON, PHYS, RECORDING_TIME, ROOT, recording_time_spec, oracle_cost_spec, oracle_regions, oracle_spec, oracle_weights, recorded
from oracle import cost as CO
from oracle import decode as D
from oracle import dynfeas, mppi, scorenet
from synthesize_pregrasp_b200 import _abi
from synthesize_pregrasp_b200.envspec import ENV_NAMES, make_spec

G = os.path.join(ROOT, "tests", "golden")
DECODE_EXPS = ["plate_20.0", "foodbox_20.0", "foodbox_50.0", "keyboard_20.0", "keyboard_50.0", "wall_laptop_20.0",
               "cardboard_20.0", "waterbottle_20.0", "keyboard_scale_20.0", "cardboard_scale_20.0", "cardboard2_20.0", "cardboard3_20.0",
               "plate_ablation_df_20.0", "plate_no_df_20.0"]


def _qrot(q, v):
    x, y, z, w = q
    u = np.array([x, y, z])
    return v + 2 * np.cross(u, np.cross(u, v) + w * v)


@pytest.mark.parametrize("exp", DECODE_EXPS)
def test_decode_reproduces_recorded_tip_positions(exp):
    """data/tip_data/<exp>_tip_poses.npy = object pose o local contact point (model_test.py:124-135)."""
    g = recorded(exp)
    spec = recording_time_spec(exp)
    reg = oracle_regions(spec, spec.dummy_states)
    carry = D.new_carry()
    from oracle.envsim import b3_multiply_transforms
    worst = worst32 = 0.0
    for j in range(2):
        pos, on, _ = D.decode_step(g["u"][j], carry, reg, [0, 0, 0], j)
        for t in range(50):
            row = j * 50 + t
            for f in range(2):
                if on[f]:
                    w = _qrot(g["poses"][row, 3:], pos[f]) + g["poses"][row, :3]
                    worst = max(worst, np.abs(w - g["tips"][row, f]).max())
                    # PyBullet's multiplyTransforms is single precision: emulated, the recording is reproduced bit for bit
                    # (plate: the mesh-region point differs from the reference's by ~1e-9, i.e. at most one float32 ulp)
                    w32 = b3_multiply_transforms(g["poses"][row, :3], g["poses"][row, 3:], pos[f])
                    worst32 = max(worst32, np.abs(w32 - g["tips"][row, f]).max())
    assert worst <= 1e-7, worst
    assert worst32 <= (2e-8 if g["env"] == "plate" else 0.0), worst32


def test_score_net_matches_reference_network():
    """fixtures made by running neurals/network.py::LargeScoreFunction on CPU (tools/make_goldens.py)."""
    w = oracle_weights()
    g = np.load(os.path.join(G, "score_net.npz"))
    for i in range(int(g["n"])):
        s = scorenet.score_logit(w, g[f"pts{i}"], g[f"df{i}"], g[f"grasp{i}"])
        assert abs(s - float(g[f"logit64_{i}"])) <= 1e-9 * max(1.0, abs(s))
        assert abs(s - float(g[f"logit32_{i}"])) <= 2e-5 * max(1.0, abs(s))     # the reference's own fp32 run


def test_mppi_update_matches_reference_lines():
    m = np.load(os.path.join(G, "mppi_update.npz"))
    u1, S = mppi.mppi_update(m["u0"], m["J"], np.transpose(m["E"], (2, 0, 1)), float(m["kappa"]), float(m["alpha"]))
    assert np.array_equal(u1, m["u1"]) and np.array_equal(S, m["S"])
    # sharded partials merge to the same update
    E = np.transpose(m["E"], (2, 0, 1))
    parts = [mppi.partials(m["J"][a:b], E[a:b], float(m["kappa"])) for a, b in ((0, 100), (100, 217), (217, 360))]
    mm, s, v, sj = mppi.merge_partials(parts, float(m["kappa"]))
    np.testing.assert_allclose(m["u0"] + float(m["alpha"]) * v / s, m["u1"], rtol=0, atol=1e-13)


def test_dyn_simple_matches_reference_function():
    d = np.load(os.path.join(G, "dyn_simple.npz"))
    got = np.array([dynfeas.simple_check_dyn_feasible(p, n) for p, n in zip(d["pts"], d["nrm"])])
    assert np.array_equal(got, d["out"])
    gotm = np.array([dynfeas.simple_check_dyn_feasible(p[:c], n[:c]) for p, n, c in zip(d["m_pts"], d["m_nrm"], d["m_count"])])
    assert np.array_equal(gotm, d["m_out"])
    assert d["out"].sum() > 100 and (~d["out"]).sum() > 100


def test_cost_geometry_against_brute_force():
    rng = np.random.RandomState(3)
    spec = make_spec("bookshelf")
    pts = rng.uniform(-0.5, 0.5, (300, 3)) * [1, 1, 0.6] + [0, 0, 0.2]
    q32 = pts.astype(np.float32).astype(np.float64)
    # exact box distance == min over the 12 triangles of each box
    def box_tris(lo, hi):
        c = np.array([[lo[0] if i & 1 == 0 else hi[0], lo[1] if i & 2 == 0 else hi[1], lo[2] if i & 4 == 0 else hi[2]] for i in range(8)])
        f = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
        return c[np.array(f)]
    tris = np.concatenate([box_tris(lo, hi) for lo, hi in spec.dist_boxes])
    brute = CO.point_triangle_distance(q32, tris)
    fast = np.min([CO.box_unsigned_distance(q32, lo, hi) for lo, hi in spec.dist_boxes], axis=0)
    np.testing.assert_allclose(fast, brute, rtol=0, atol=1e-12)
    # crop keeps duplicates per box and the back-transform returns the cloud
    pose = np.r_[0.01, -0.02, 0.21, 0.5, -0.5, 0.5, 0.5]
    p32, df, idx = CO.cost_inputs(pose[:3], pose[3:], spec.cloud, spec.deocclude_boxes, spec.dist_boxes)
    assert len(idx) > 0 and np.abs(p32 - spec.cloud[idx].astype(np.float32)).max() < 1e-6
    assert (df >= 0).all()


@pytest.mark.parametrize("exp", sorted(PHYS))
def test_physics_oracle_against_recorded_trajectories(exp):
    """The only physics ground truth: data/object_poses/<exp>_object_poses.npy (PyBullet, recorded by model_test.py)."""
    from oracle.envsim import EnvSim
    g = recorded(exp)
    osd = oracle_spec(recording_time_spec(exp))
    sim = EnvSim(osd, g["max_force"])
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    rows, tol = PHYS[exp]
    err = np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"]))   # q and -q are one rotation
    assert err[0].max() <= 1e-12      # reset + pre-simulate substeps: exact
    assert err[list(rows)].max() <= tol, err[list(rows)].max(axis=1)
    # whole 100-substep horizon stays physically close (contact-rich motion is chaotic)
    assert err[:, :3].max() <= 5e-2
    if exp == "waterbottle_20.0":
        assert err[:33].max() <= 1.5e-8
    if exp == "keyboard_20.0":
        assert err[:24].max() <= 1e-12 and err.max() <= 1e-11          # the whole horizon at rounding level
    if exp == "foodbox_20.0":
        assert err[:13].max() <= 1e-12 and err[:18].max() <= 1.5e-8
    if exp == "plate_20.0":
        assert err[:7].max() <= 1e-10
        assert err[:62].max() <= 1e-6 and err.max() <= 2e-5      # env step 2: within 1e-6 over 11 more substeps, 1e-5 over the whole horizon
    if exp == "cardboard_20.0":
        assert err[:70].max() <= 1e-12 and err.max() <= 1e-11          # the whole horizon at rounding level
    if exp in HORIZON:
        assert err.max() <= HORIZON[exp], err.max()


@pytest.mark.parametrize("exp", ["cardboard2_20.0", "cardboard3_20.0", "plate_no_df_20.0", "plate_ablation_df_20.0"])
def test_held_out_recordings_second_env_step_is_a_pair_order(exp):
    """The two held-out recordings whose thumb stays on the plate through both env steps are exact over env step 1 with the frozen
    model and leave it in the first substep of env step 2 (7e-6 / 3e-6).  One thing explains both: there the re-created object proxy
    finds the thumb BEFORE the floor (oracle knob find_order_late = 3429) -- with it all 100 rows are reproduced at 2e-12 / 2e-11.  The
    model does not adopt it as a rule: keyboard_scale_20.0 (thumb on in both env steps as well) needs the floor first, i.e. the order
    follows Bullet's broadphase tree (re-sorted by node address), which can be fitted per recording but not derived."""
    from oracle.envsim import EnvSim
    g = recorded(exp)
    sim = EnvSim(oracle_spec(recording_time_spec(exp)), g["max_force"], params=dict(find_order_late=3429))
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    err = np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max()
    # (plate_no_df_20.0: thumb new in env step 2, 2.3e-8; plate_ablation_df_20.0: 3.9e-11 instead of the frozen model's 1.3e-5)
    assert err <= {"plate_no_df_20.0": 5e-8, "plate_ablation_df_20.0": 1e-9}.get(exp, 1e-10), err


@pytest.mark.parametrize("exp", sorted(PHYS))
def test_product_kernel_logic_on_host_follows_recordings(exp):
    """csrc/pg_rollout.cuh (the per-thread code of the CUDA kernel) compiled for the host, replaying the reference's recorded
    controls: the same rows and bounds against the PyBullet recording as the oracle above (the GPU test repeats it on the device)."""
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio",
                           "-o", so, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    H = C.CDLL(so)
    g = recorded(exp)
    spec = recording_time_spec(exp)
    cs, keep = _abi.make_c_spec(spec)
    N = 2
    u = np.ascontiguousarray(g["u"], np.float64)
    pose = np.zeros((1, N, 7)); fins = np.zeros((1, N, 2, 3)); metric = np.zeros(1); trace = np.zeros((1, N * 50, 7))
    path = np.zeros(N, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert H.hostsim_rollout(C.byref(cs), vp(u), None, 1, N, vp(path), C.c_double(g["max_force"]), 0, vp(pose), vp(fins), vp(metric),
                             vp(trace), None) == 0
    tr = trace[0]
    e = np.minimum(np.abs(tr - g["poses"]), np.abs(tr * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max(axis=1)
    rows, tol = PHYS[exp]
    assert e[0] <= 1e-12, e[0]
    assert e[list(rows)].max() <= tol, (e[list(rows)].max(), tol)
    if exp in HORIZON:
        assert e.max() <= HORIZON[exp], e.max()


def test_physics_oracle_parked_fingertip_meets_idle_ones():
    """wall_laptop_20.0 parks its thumb at (100, 100, 100) in the first env step -- where the env's idle fingertips were created.  In
    PyBullet they collide; the manifolds this creates and releases (world step 9) re-order the object's rows.  With that interaction
    switched on the oracle follows the recording three rows further (7-9) at rounding level; it is off by default because the CUDA
    kernel carries the idle fingertips only as entries of the dispatcher array."""
    from oracle.envsim import EnvSim
    g = recorded("wall_laptop_20.0")
    sim = EnvSim(oracle_spec(make_spec(g["env"])), g["max_force"], params=dict(tip_idle_collide=1))
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    err = np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"]))
    assert err[:10].max() <= 1e-12, err[:10].max(axis=1)


def test_idle_fingertips_timeline():
    """The three idle fingertips alone (every env creates them at the same point): their pairs' fat AABBs separate during world
    steps 20 (outer pair) and 39 (both inner pairs) -- the constants the CUDA kernel uses (pg_physics.cuh idle_pair_alive) -- for
    every fingertip size and friction the envs use."""
    import ctypes
    from oracle import envsim as OE
    L = OE.lib()
    dp = ctypes.POINTER(ctypes.c_double)
    for hz, fr in ((0.01, 5.0), (0.011, 100.0), (0.012, 100.0), (0.011, 150.0)):
        h = L.pgo_create()
        keep = []
        for f in range(3):
            arrs = [np.array([0.01, 0.01, hz]), np.zeros((0, 3)), np.array([100.0, 100.0, 100.0]), np.array([0, 0, 0, 1.0])]
            keep.append(arrs)
            L.pgo_add_body(h, 0, arrs[0].ctypes.data_as(dp), arrs[1].ctypes.data_as(dp), 0, 0.01, arrs[2].ctypes.data_as(dp),
                           arrs[3].ctypes.data_as(dp), fr, 0, 0.001)
        L.pgo_set_roles(h, -1, -1, -1)
        counts = []
        for s in range(1, 45):
            L.pgo_step(h)
            counts.append(L.pgo_num_manifolds(h))
        L.pgo_destroy(h)
        assert counts[:19] == [3] * 19 and counts[19:38] == [2] * 19 and counts[38:] == [0] * 6, counts


def test_physics_oracle_resting_contact_exact():
    """keyboard_scale_20.0 was recorded with another contact state (thumb on region 2; the committed dummy_states.npy says 9): its
    thumb hovers 1.2e-8 above the top face for the whole first env step (the committed env's overlaps by 1.1e-8), so PyBullet has no
    fingertip contact there.  (i) With the committed state and the fingertip pairs switched off for that env step the restatement
    reproduces the recording at rounding level over 52 substeps of resting contact on the floor (object vs compound floor: box-box,
    persistent manifold, per-island solve, slop, integration); (ii) with the recording-time state (helpers.RECORDING_TIME) the decoded
    fingertip is the recorded one bit for bit and the WHOLE recording is reproduced at rounding level with nothing switched off."""
    from oracle.envsim import EnvSim, b3_multiply_transforms
    g = recorded("keyboard_scale_20.0")
    rt = recording_time_spec("keyboard_scale_20.0")
    pos, on, _ = D.decode_step(g["u"][0], D.new_carry(), oracle_regions(rt, rt.dummy_states), [0, 0, 0], 0)
    assert on[0] and not on[1]
    assert np.array_equal(b3_multiply_transforms(g["poses"][0, :3], g["poses"][0, 3:], pos[0]), g["tips"][0, 0])
    sim = EnvSim(oracle_spec(rt), g["max_force"])
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    assert np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max() <= 1e-14
    # the recorded thumb does hover: its bottom face against the recorded top face, rows 0-49
    R22 = 1.0                                                # tilt 1e-7 rad: cos = 1 to 1e-14
    gap = (g["tips"][:50, 0, 2] - 0.01) - (g["poses"][:50, 2] + 0.03 * R22)
    assert 0 < gap.min() and gap.max() < 3e-8
    spec = make_spec(g["env"])
    sim = EnvSim(oracle_spec(spec), g["max_force"], params=dict(tip_skip=3, tip_skip_n=48))
    sim.reset()
    P = np.concatenate([sim.step(g["u"][j], [0, 0, 0])["poses"] for j in range(2)])
    err = np.minimum(np.abs(P - g["poses"]), np.abs(P * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"]))
    assert err[:52].max() <= 2e-15, err[:52].max(axis=1)


def test_abi_library_loads_and_exports_header_symbols(built):
    L = _abi.lib()
    hdr = open(os.path.join(ROOT, "include", "pregrasp_b200.h")).read()
    declared = set(re.findall(r"\b(pg_[a-z_]+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for f in declared:
        assert hasattr(L, f), f
    assert b"sm_100a" in L.pg_version()
    # host-only entry point works without a GPU
    u = np.zeros((2, 48))
    P = np.zeros((2, 3 + 96))
    P[0, :3] = [1.0, 2.0, 0.0]; P[1, :3] = [1.2, 3.0, 0.0]
    P[0, 3:] = 1.0; P[1, 3:] = 2.0
    from synthesize_pregrasp_b200.engine import mppi_merge_apply
    out = mppi_merge_apply(P, u, kappa=5.0, alpha=1.0)
    w1 = np.exp(-5 * 0.2)
    np.testing.assert_allclose(out, (1.0 + 2.0 * w1) / (2.0 + 3.0 * w1), rtol=1e-14)


def test_sass_is_sm100_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_every_env_spec_builds_and_c_spec_roundtrips():
    for name in ENV_NAMES:
        spec = make_spec(name)
        assert spec.cloud.shape == (2048, 3)
        cs, keep = _abi.make_c_spec(spec)
        assert cs.n_static == len(spec.statics) and cs.n_cloud == 2048
        assert spec.substeps_per_rollout in (203, 403)


def test_product_kernel_logic_on_host_matches_oracle():
    """csrc/pg_rollout.cuh (the per-thread code of the CUDA kernel) compiled for the host vs the independent oracle."""
    from oracle.envsim import EnvSim
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio",
                           "-o", so, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    H = C.CDLL(so)
    rng = np.random.RandomState(7)
    for env, F in (("keyboard", 20.0), ("bookshelf", 50.0), ("laptop", 20.0), ("plate", 20.0), ("waterbottle", 20.0),
                   ("foodbox", 20.0), ("cardboard", 20.0)):
        spec = make_spec(env)
        cs, keep = _abi.make_c_spec(spec)
        u = 0.8 * rng.randn(2, 48)
        N = 2
        pose = np.zeros((1, N, 7)); fins = np.zeros((1, N, 2, 3)); metric = np.zeros(1); trace = np.zeros((1, N * 50, 7))
        path = np.zeros(N, np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        assert H.hostsim_rollout(C.byref(cs), vp(u), None, 1, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric),
                                 vp(trace), None) == 0
        sim = EnvSim(oracle_spec(spec), F)
        sim.reset()
        rs = [sim.step(u[j], [0, 0, 0], train=True, n_steps=N) for j in range(N)]
        P = np.concatenate([r["poses"] for r in rs])
        d = np.abs(trace[0] - P).max(axis=1)
        # identical algorithm, independently written: agreement to rounding until chaos amplifies it
        assert d[:20].max() <= 1e-10, (env, d[:20].max())
        assert np.median(d) <= 1e-8, (env, np.median(d))
        for j in range(N):
            np.testing.assert_allclose(fins[0, j], rs[j]["cur_pos"], rtol=0, atol=1e-12)


def test_csg_path_selection_matches_bruteforce():
    """generate_csg.select_paths (reference: neurals/scripts/generate_csg.py:136-180) vs an independent enumeration."""
    import itertools
    from synthesize_pregrasp_b200 import generate_csg as G
    rng = np.random.RandomState(3)
    regions = [1, 2, 3, 5, 8]
    nodes = G.all_ordered_nodes(regions)
    assert len(nodes) == 20 and nodes[0] == (1, 2) and nodes[-1] == (8, 5)
    all_scores = rng.randn(len(nodes))
    init_idx = [i for i, n in enumerate(nodes) if n[0] in (1, 2, 3) and n[1] in (1, 2, 3)]
    init_nodes = [nodes[i] for i in init_idx]
    init_scores = rng.randn(len(init_nodes))
    table, ids, sorted_paths, sorted_scores = G.select_paths(init_nodes, init_scores, nodes, all_scores, k=2, top=10)
    brute = []
    for (a, sa), (b, sb) in itertools.product(zip(init_nodes, init_scores), zip(nodes, all_scores)):
        if b[0] == a[0] or b[1] == a[1]:
            brute.append((sa + sb, a, b))
    brute.sort(key=lambda t: t[0])
    assert len(brute) == len(sorted_paths)
    np.testing.assert_allclose(sorted_scores, [t[0] for t in brute], rtol=0, atol=1e-15)
    for p, t in zip(sorted_paths[:10], brute[:10]):
        assert tuple(p[0]) == t[1] and tuple(p[1]) == t[2]
    assert ids.shape == (10, 2)
    for row, p in zip(ids, sorted_paths[:10]):                 # the id table indexes back to the chosen nodes
        assert [tuple(table[i]) for i in row] == [tuple(n) for n in p]
    # three-node paths: every consecutive pair keeps one finger's region
    _, _, sp3, _ = G.select_paths(init_nodes, init_scores, nodes, all_scores, k=3, top=5)
    for p in sp3[:50]:
        for a, b in zip(p[:-1], p[1:]):
            assert a[0] == b[0] or a[1] == b[1]


def test_kernel_logic_is_clean_under_address_and_ub_sanitizers():
    """compute-sanitizer is not available on the GPU pool, so the per-thread kernel code (box-box, GJK, polytope pools,
    explicit stacks, manifold arrays, solver rows) is compiled for the host with -fsanitize=address,undefined and run on
    random rollouts of all seven envs."""
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined", "-shared", "-include", "cstdio", "-o", so,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    asan = subprocess.check_output(["g++", "-print-file-name=libasan.so"]).decode().strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "hostsim", "asan_driver.py"), so], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count(" ok") == 7, out.stdout


def test_qp_solver_logic_on_host_matches_independent_solve():
    """csrc/pg_qp.cuh (non-negative least squares in generator space, the code of the QP kernels) compiled for the host vs
    SLSQP on the reference's (Q, G, h) formulation: random triples, closure-like triples, degenerate frames (n = z),
    unnormalised / axis-aligned normals."""
    from oracle import dynfeas as DF
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio",
                           "-o", so, os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    H = C.CDLL(so)
    H.hostsim_qp.restype = C.c_double
    rng = np.random.RandomState(5)
    n_feasible = 0
    for t in range(250):
        p = rng.uniform(-0.2, 0.2, (3, 3))
        nrm = rng.randn(3, 3)
        nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        if t % 3 == 0:
            nrm = p - p.mean(0)
            nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        if t % 7 == 0:
            nrm[0] = [0, 0, 1.0]
        if t % 11 == 0:
            nrm = np.round(nrm * 2) / 2
            nrm[np.linalg.norm(nrm, axis=1) == 0] = [1, 0, 0]
        it = C.c_int(0)
        got = H.hostsim_qp(p.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), C.c_double(5.0), C.c_double(1.0), C.byref(it))
        ref = DF.qp_objective(p, nrm, normalise=False)
        assert abs(got - ref) <= 1e-6 * max(1.0, ref) + 2e-9, (t, got, ref, it.value)
        n_feasible += got <= 1e-4
    assert 50 < n_feasible < 250            # both outcomes are exercised


def test_model_optimize_driver_arguments_and_paths():
    """model_optimize.py:36-83: same argument names and defaults, path list from paths_id.npy under --validate."""
    from synthesize_pregrasp_b200 import model_optimize as MO
    a = MO.build_parser().parse_args(["--env", "plate", "--max_force", "20", "--has_distance_field", "--name_score", "x"])
    assert (a.iters, a.steps, a.runs, a.validate, a.exp_name) == (20, 2, 3, False, "u_opt_0-10_tp4")
    kw = MO.optimiser_kwargs(a, 20.0)
    assert kw["Num_processes"] * kw["Traj_per_process"] == 360 and kw["sigma"] == 0.8 and kw["max_forces"] == 20.0
    a.validate = True
    assert MO.optimiser_kwargs(a, 20.0)["Traj_per_process"] == 15
    assert MO.path_list("plate", 2, False, 10) == [[0, 0, 0]]
    paths = MO.path_list("plate", 2, True, 10)
    spec = make_spec("plate")
    assert len(paths) == min(10, len(spec.paths_id)) and all(0 <= s < len(spec.csg_states) for p in paths for s in p)
    assert MO.build_parser().parse_args([]).max_force == -1 and MO.MAX_FORCE_DEFAULT == 80.0


@pytest.mark.parametrize("exp", ["plate_20.0", "foodbox_20.0", "waterbottle_20.0", "wall_laptop_20.0"])
def test_tip_records_reproduce_recorded_tip_data(exp):
    """model_test.py:115-135 recording format: object pose o local contact point + nearest-cloud-point normal for the two active
    fingers, 100 for the three absent ones -- data/tip_data/<exp>_tip_poses.npy[:100] (tests/golden/recorded_tip_records.npz)."""
    from synthesize_pregrasp_b200.model_test import tip_records
    g = recorded(exp)
    raw = np.load(os.path.join(ROOT, "tests", "golden", "recorded_tip_records.npz"))[exp.replace(".", "p")]
    spec = make_spec(g["env"])
    reg = oracle_regions(spec, spec.dummy_states)
    carry = D.new_carry()
    local, on = [], []
    for j in range(2):
        pos, o, _ = D.decode_step(g["u"][j], carry, reg, [0, 0, 0], j)
        local.append(pos.copy()); on.append(o.copy())
    got = tip_records(g["poses"], np.stack(local), spec.cloud, spec.cloud_normals)
    assert got.shape == raw.shape == (100, 5, 6) and np.all(got[:, 2:] == 100.0) and np.all(raw[:, 2:] == 100.0)
    for j in range(2):
        rows = slice(50 * j, 50 * j + 50)
        for f in range(2):
            # multiplyTransforms emulated in float32: bit-identical to the recording, fingers that are off (local point
            # (100,100,100)) included; the plate's mesh-region points differ from the reference's by ~1e-9, i.e. <= 1 float32 ulp
            if exp == "plate_20.0":
                assert np.abs(got[rows, f, :3] - raw[rows, f, :3]).max() <= (2e-8 if on[j][f] else 1e-5)
            else:
                np.testing.assert_array_equal(got[rows, f, :3], raw[rows, f, :3])
            if on[j][f]:
                assert np.abs(got[rows, f, 3:] - raw[rows, f, 3:]).max() <= (2e-7 if exp == "plate_20.0" else 0.0)


def test_single_precision_multiply_transforms_native_equals_numpy_statement():
    """oracle/envsim.py b3_multiply_transforms (numpy float32, the documented definition) == the C++ copy EnvSim calls."""
    from oracle.envsim import EnvSim, b3_multiply_transforms, b3_multiply_transforms_native
    sim = EnvSim(oracle_spec(make_spec("foodbox")), 20.0)
    rng = np.random.RandomState(3)
    for _ in range(2000):
        p = rng.randn(3); q = rng.randn(4); q /= np.linalg.norm(q) * rng.uniform(0.999999, 1.000001)
        v = rng.randn(3) * rng.choice([1e-3, 0.1, 1.0, 100.0])
        np.testing.assert_array_equal(b3_multiply_transforms(p, q, v), b3_multiply_transforms_native(sim.L, p, q, v))


def test_roofline_traffic_record_matches_this_build(built):
    """bench.py reports roofline.traffic only for the build the ncu capture was taken from: matched by the hash of the library's
    sources or -- when only a header that the captured kernel does not compile has changed -- by the hash of that kernel's SASS."""
    import json
    import shutil
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    if shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    sha = bench._kernel_sass_sha16(bench.BOX_ROLLOUT_KERNEL)
    assert sha is not None and len(sha) == 16
    assert bench._kernel_sass_sha16("_ZN2pg7no_suchEv") is None                      # unknown kernel -> no match, traffic null
    for wl, want in rec.get("kernel_sass_sha16", {}).items():
        assert (bench._ncu_traffic(wl) is not None) == (rec["src_sha16"] == bench._src_sha16() or want == sha)
    assert bench._ncu_traffic("no_such_workload") is None
