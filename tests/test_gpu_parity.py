"""GPU parity tests: every call goes through the C ABI (libpregrasp_b200.so); the oracle is the checker."""
import os

import numpy as np
import pytest

from helpers import ROOT, oracle_cost_spec, oracle_spec, oracle_weights, recorded

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def EN():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from synthesize_pregrasp_b200 import engine
    return engine


def _oracle_trace(spec, u, noise_row, F, train=False):
    from oracle.envsim import EnvSim
    sim = EnvSim(oracle_spec(spec), F)
    sim.reset()
    rs = []
    for j in range(u.shape[0]):
        a = u[j] + (noise_row[j].astype(np.float64) if noise_row is not None else 0.0)
        rs.append(sim.step(a, [0, 0, 0], train=train, n_steps=u.shape[0]))
    return rs


def _recording_time_spec(exp, env):
    """The env as it was when `exp` was recorded (tests/test_oracle_cpu.py RECORDING_TIME: constants identified from the data)."""
    from helpers import recording_time_spec
    return recording_time_spec(exp, env)


@pytest.mark.parametrize("exp", ["plate_20.0", "waterbottle_20.0", "foodbox_20.0", "keyboard_20.0", "keyboard_scale_20.0", "wall_laptop_20.0", "cardboard_20.0",
                                 # held-out recordings (tests/helpers.py HELD_OUT): replayed with the frozen model
                                 "cardboard_scale_20.0", "cardboard2_20.0", "cardboard3_20.0", "foodbox_ablation_df_20.0",
                                 "plate_ablation_df_20.0", "plate_no_df_20.0"])
def test_physics_kernel_on_recorded_controls(EN, exp):
    """K=1 replay of the reference's recorded u: kernel vs oracle (tight) and vs the PyBullet recording over the SAME rows and
    tolerances the oracle itself is held to (tests/test_oracle_cpu.py PHYS)."""
    from helpers import PHYS
    g = recorded(exp)
    spec = _recording_time_spec(exp, g["env"])
    eng = EN.RolloutEngine(spec, device=0)
    out = eng.physics(g["u"], None, [0, 0], g["max_force"], K=1, want_trace=True)
    tr = out["trace"][0].cpu().numpy()
    P = np.concatenate([r["poses"] for r in _oracle_trace(spec, g["u"], None, g["max_force"])])
    d = np.abs(tr - P).max(axis=1)
    rows, tol = PHYS[exp]
    assert d[:20].max() <= 1e-9, d[:20].max()              # same algorithm: rounding-level agreement early on
    # ... and over every row on which the oracle is pinned to PyBullet; past those rows both follow their own rounding (a tilting
    # box held by two fingertips and one floor corner amplifies 1e-12 to 1e-3 within ten substeps -- foodbox after row 29)
    assert d[list(rows)].max() <= max(1e-9, 0.1 * tol), (d[list(rows)].max(), tol)
    if exp != "foodbox_20.0":
        assert np.median(d) <= 1e-7, np.median(d)
    e = np.minimum(np.abs(tr - g["poses"]), np.abs(tr * np.r_[1, 1, 1, -1, -1, -1, -1] - g["poses"])).max(axis=1)   # q == -q
    # the kernel contracts a*b+c into FMAs (box objects) where PyBullet and the oracle round twice: rounding-level slack on top
    assert e[0] <= 1e-11, e[0]
    assert e[list(rows)].max() <= 4 * tol + 1e-10, (e[list(rows)].max(), tol)
    from helpers import HORIZON
    if exp in HORIZON:                                     # followed to the end of the recording: the north-star bound is 1e-4
        assert e.max() <= min(1e-4, 5 * HORIZON[exp]), e.max()
    print(f"{exp}: kernel vs PyBullet recording, rows {rows.start}-{rows.stop - 1}: max {e[list(rows)].max():.2e} (oracle bound {tol:.0e}); whole horizon {e.max():.2e}")


@pytest.mark.parametrize("env,F", [("plate", 20.0), ("waterbottle", 20.0), ("foodbox", 20.0), ("bookshelf", 50.0), ("keyboard", 20.0)])
def test_physics_kernel_random_batch(EN, env, F):
    """Seeded noise batch, tolerance 1e-4 on the object pose at the cost-evaluation instants (north_star)."""
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(env)
    eng = EN.RolloutEngine(spec, device=0)
    K, N = 48, 2
    gen = torch.Generator(device="cuda:0").manual_seed(1234)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((N, 48))
    out = eng.physics(u, noise, [0, 0], F, want_trace=True)
    pose = out["pose"].cpu().numpy(); fins = out["fins"].cpu().numpy(); metric = out["metric"].cpu().numpy()
    nz = noise.cpu().numpy()
    ok = 0
    for k in range(K):
        rs = _oracle_trace(spec, u, nz[k], F, train=True)
        np.testing.assert_allclose(fins[k], np.stack([r["cur_pos"] for r in rs]), rtol=0, atol=1e-12)
        ref = np.stack([r["final_pose"] for r in rs])
        err = np.abs(pose[k] - ref).max()
        errq = np.abs(pose[k] * np.r_[1, 1, 1, -1, -1, -1, -1] - ref).max()
        ok += min(err, errq) <= 1e-4
        # the first 20 substeps must agree tightly for every rollout (checked on the first 24 of the batch)
        if k < 24:
            assert np.abs(out["trace"][k, :20].cpu().numpy() - np.concatenate([r["poses"] for r in rs])[:20]).max() <= 1e-8
    # contact-rich rollouts are chaotic: a rounding-level difference can flip a contact event late in the
    # 200-400 substeps; require the bulk to agree within 1e-4 and report the fraction
    print(f"{env}: {ok}/{K} rollouts within 1e-4 of the oracle at both cost instants")
    # measured agreement (160 rollouts on the GPU, 96 with the host build of the kernel code): 85-99 %, the toppling bottle
    # (403 substeps, polytope contacts) 73-79 %; required of these 48: 70 % and 55 %, so that the sample's own scatter
    # (binomial, sd ~2.5 rollouts) cannot fail the test
    assert ok >= int({"waterbottle": 0.60, "bookshelf": 0.80}.get(env, 0.90) * K), ok


@pytest.mark.parametrize("env", ["plate", "waterbottle", "foodbox", "bookshelf", "keyboard", "laptop", "cardboard"])
def test_cost_kernel_matches_oracle(EN, env):
    from oracle import cost as CO
    from oracle import scorenet as SN
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(env)
    eng = EN.RolloutEngine(spec, device=0)
    rng = np.random.RandomState(5)
    B = 12
    pose = np.zeros((B, 7)); fins = np.zeros((B, 2, 3))
    for b in range(B):
        q = np.r_[spec.obj_quat] + 0.15 * rng.randn(4)
        pose[b] = np.r_[np.r_[spec.obj_pos] + 0.03 * rng.randn(3) + [0, 0, 0.02], q / np.linalg.norm(q)]
        for f in range(2):
            fins[b, f] = spec.cloud[rng.randint(2048)] + 0.01 * rng.randn(3) if rng.rand() < 0.7 else [100, 100, 100]
    logit = eng.cost(torch.as_tensor(pose, device="cuda:0"), torch.as_tensor(fins, device="cuda:0")).cpu().numpy()
    cs, w = oracle_cost_spec(spec), oracle_weights()
    for b in range(B):
        pts, df, _ = CO.cost_inputs(pose[b, :3], pose[b, 3:], cs["cloud"], cs["deocclude_boxes"], cs["dist_boxes"], cs["dist_tris"])
        grasp = np.concatenate([CO.project_to_pcd(fins[b, f], cs["cloud"], cs["clearance_h"], cs["walls"], cs["wall_size"], cs["strict_z"]) for f in range(2)])
        ref = SN.score_logit(w, pts, df, grasp) if len(pts) else float("nan")
        if np.isnan(ref):
            assert np.isnan(logit[b])
        else:
            # fp32 network arithmetic vs the fp64 evaluation of the same fp32 weights
            assert abs(logit[b] - ref) <= 2e-5 * max(1.0, abs(ref)), (b, logit[b], ref)


def test_score_entry_point_matches_reference_fixture(EN):
    from synthesize_pregrasp_b200.envspec import make_spec
    eng = EN.RolloutEngine(make_spec("foodbox"), device=0)
    g = np.load(os.path.join(G, "score_net.npz"))
    for i in range(int(g["n"])):
        pts = torch.as_tensor(g[f"pts{i}"][None], device="cuda:0"); df = torch.as_tensor(g[f"df{i}"][None], device="cuda:0")
        gr = torch.as_tensor(g[f"grasp{i}"][None], device="cuda:0")
        s = float(eng.score(pts, df, gr).item())
        ref = float(g[f"logit64_{i}"])
        assert abs(s - ref) <= 2e-5 * max(1.0, abs(ref)), (i, s, ref)
    # ragged batch through the mask: first 977 points valid
    pts = torch.as_tensor(g["pts0"][None], device="cuda:0"); df = torch.as_tensor(g["df0"][None], device="cuda:0")
    mask = torch.zeros((1, 2048), dtype=torch.uint8, device="cuda:0"); mask[:, :977] = 1
    s = float(eng.score(pts, df, torch.as_tensor(g["grasp0"][None], device="cuda:0"), mask).item())
    from oracle import scorenet as SN
    ref = SN.score_logit(oracle_weights(), g["pts0"][:977], g["df0"][:977], g["grasp0"])
    assert abs(s - ref) <= 2e-5 * max(1.0, abs(ref))


def test_end_to_end_cost_and_update_selection(EN):
    """J through pg_rollout (device) and pg_rollout_host (host buffers) vs the oracle; same MPPI update."""
    from oracle import mppi
    from oracle.rollout import rollout as oracle_rollout
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec("foodbox")
    eng = EN.RolloutEngine(spec, device=0)
    K, N = 16, 2
    gen = torch.Generator(device="cuda:0").manual_seed(99)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((N, 48))
    J, metric = eng.rollout(u, noise, [0, 0], 20.0, want_metric=True)
    Jh, mh = eng.rollout_host(u, noise.cpu().numpy(), [0, 0], 20.0, want_metric=True)
    J = J.cpu().numpy(); metric = metric.cpu().numpy()
    assert np.array_equal(J, Jh) and np.array_equal(metric, mh)       # same kernels, same bits
    ref = [oracle_rollout(oracle_spec(spec), oracle_cost_spec(spec), oracle_weights(), u, noise[k].cpu().numpy(), [0, 0, 0], 20.0) for k in range(K)]
    Jr = np.array([r["J"] for r in ref]); mr = np.array([r["metric"] for r in ref])
    close = np.abs(J - Jr) <= 1e-3 * np.maximum(1.0, np.abs(Jr))
    print("rollout costs within 1e-3 rel:", int(close.sum()), "of", K, "max |dJ|", np.abs(J - Jr)[close].max())
    assert close.sum() >= int(0.75 * K)
    assert np.abs(metric - mr)[close].max() <= 1e-4
    # update on the oracle's costs vs update on the kernel's costs (restricted to agreeing rollouts it is identical)
    u_t = torch.zeros((N, 48), dtype=torch.float64, device="cuda:0")
    part = EN.mppi_update(torch.as_tensor(J, device="cuda:0"), noise, u_t, 5.0, 1.0, want_partial=True)
    un, S = mppi.mppi_update(u, J, noise.cpu().numpy(), 5.0, 1.0)
    np.testing.assert_allclose(u_t.cpu().numpy(), un, rtol=0, atol=1e-12)
    assert abs(float(part[0]) - J.min()) == 0.0
    assert int(np.argmax(S)) == int(np.argmin(J))


@pytest.mark.parametrize("env,F", [("plate", 20.0), ("bookshelf", 50.0)])
def test_update_selection_kernel_vs_oracle(EN, env, F):
    """north_star: "select the same MPPI update".  At the reference's own K = 12 x 30 = 360 samples (model_optimize.py:77) the
    update computed from the KERNEL's costs is compared with the update computed from the ORACLE's costs on identical noise
    (stoch_traj_opt.py:195-204): S = exp(-5 (J - min J)) is razor sharp, so one diverged low-cost rollout would own the update."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import mppi
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(env)
    eng = EN.RolloutEngine(spec, device=0)
    K, N = 360, 2
    gen = torch.Generator(device="cuda:0").manual_seed(12367134)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((N, 48))
    J = eng.rollout(u, noise, [0, 0], F).cpu().numpy()
    nz = noise.cpu().numpy()
    with ProcessPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        Jr = np.array(list(ex.map(_oracle_J, [(env, F, nz[k]) for k in range(K)], chunksize=8)))
    un_k, S_k = mppi.mppi_update(u, J, nz, 5.0, 1.0)
    un_o, S_o = mppi.mppi_update(u, Jr, nz, 5.0, 1.0)
    close = np.abs(J - Jr) <= 1e-3 * np.maximum(1.0, np.abs(Jr))
    du = np.abs(un_k - un_o).max()
    top_k, top_o = np.argsort(-S_k)[:5], np.argsort(-S_o)[:5]
    print(f"{env}: K={K} costs within 1e-3 rel: {int(close.sum())}; arg-max weight kernel {top_k[0]} (S={S_k[top_k[0]]:.4f}) oracle {top_o[0]} "
          f"(S={S_o[top_o[0]]:.4f}); top-5 kernel {top_k.tolist()} oracle {top_o.tolist()}; |du|_inf = {du:.3e}")
    assert close.sum() >= int(0.85 * K)
    assert top_k[0] == top_o[0], "the rollout that owns the update differs"
    # the weights are exp(-5 dJ): a cost difference of 1e-3 (the fp32 network's rounding, x100) moves a weight by 0.5 %
    assert du <= 2e-2, du


def _oracle_J(args):
    from oracle.rollout import rollout as oracle_rollout
    from synthesize_pregrasp_b200.envspec import make_spec
    env, F, nz = args
    spec = make_spec(env)
    return oracle_rollout(oracle_spec(spec), oracle_cost_spec(spec), oracle_weights(), np.zeros((2, 48)), nz, [0, 0, 0], F)["J"]


@pytest.mark.parametrize("env,F", [("bookshelf", 50.0), ("foodbox", 20.0), ("plate", 20.0)])
def test_device_kernel_matches_host_build_of_the_same_code(EN, env, F):
    """Deterministic check of the DEVICE port: csrc/pg_rollout.cuh compiled for the host (tests/hostsim, no FMA contraction,
    glibc math) against the CUDA build of the very same code, per control substep.  Differences can only come from FMA
    contraction (box kernel) and libm (sin / cos / tanh / atan2), i.e. rounding level until a contact event amplifies them."""
    import ctypes as C
    import subprocess
    from synthesize_pregrasp_b200 import _abi
    from synthesize_pregrasp_b200.envspec import make_spec
    so = os.path.join(ROOT, "tests", "hostsim", "libpg_hostsim.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-include", "cstdio", "-o", so,
                           os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")])
    H = C.CDLL(so)
    spec = make_spec(env)
    cs, keep = _abi.make_c_spec(spec)
    K, N = 32, 2
    gen = torch.Generator(device="cuda:0").manual_seed(77)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((N, 48))
    out = EN.RolloutEngine(spec, device=0).physics(u, noise, [0, 0], F, want_trace=True)
    tr_d = out["trace"].cpu().numpy().reshape(K, N * 50, 7)
    nz = np.ascontiguousarray(noise.cpu().numpy())
    pose = np.zeros((K, N, 7)); fins = np.zeros((K, N, 2, 3)); metric = np.zeros(K); trace = np.zeros((K, N * 50, 7)); path = np.zeros(N, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert H.hostsim_rollout(C.byref(cs), vp(u), vp(nz), K, N, vp(path), C.c_double(F), 1, vp(pose), vp(fins), vp(metric), vp(trace), None) == 0
    d = np.abs(tr_d - trace).max(axis=2)                    # [K, 100] per control substep
    np.testing.assert_allclose(out["fins"].cpu().numpy(), fins, rtol=0, atol=1e-13)  # decode: tanh / sqrt / sin of the two math libraries
    assert d[:, :20].max() <= 1e-10, d[:, :20].max()        # every rollout, first 20 substeps: rounding level
    whole = (d.max(axis=1) <= 1e-8).mean()
    print(f"{env}: device vs host build, first 20 substeps max {d[:, :20].max():.1e}; {whole * 100:.0f} % of rollouts within 1e-8 over all 100 control substeps")
    assert whole >= 0.6


@pytest.mark.parametrize("exp", ["foodbox_20.0", "plate_20.0"])
def test_env_step_single_matches_batch_trace(EN, exp):
    """pg_env_step_single + PgCarry: stepping a recorded u one env step per call (state kept on the GPU between calls) gives the
    very trace, poses, metric and costs of the batch entry points -- the same per-rollout code runs in both."""
    from synthesize_pregrasp_b200.envs import envs_dict
    from synthesize_pregrasp_b200.envspec import make_spec
    g = recorded(exp)
    spec = make_spec(g["env"])
    eng = EN.RolloutEngine(spec, device=0)
    F = g["max_force"]
    ref = eng.physics(g["u"], None, [0, 0], F, K=1, want_trace=True)
    J, metric = eng.rollout(g["u"], None, [0, 0], F, K=1, want_metric=True)
    ep = EN.Episode(eng)
    for rep in range(2):                                   # a second episode after reset() starts from the same state
        pose0 = ep.reset()
        rewards, traces = [], []
        for j in range(2):
            r, m, pose, tr = ep.step(g["u"][j], 0, F, last_step=(j == 1), want_trace=True)
            rewards.append(r); traces.append(tr)
            assert np.array_equal(pose, ref["pose"][0, j].cpu().numpy())
        assert np.array_equal(np.concatenate(traces), ref["trace"][0].cpu().numpy())
        assert m == float(metric.item())
        assert abs(-(rewards[0] + rewards[1]) - float(J.item())) <= 1e-9 * max(1.0, abs(float(J.item())))
    assert abs(pose0[2] - spec.obj_pos[2]) < 5e-3
    # the gym-style shim of the reference env classes is one such call per step()
    env = envs_dict[{"foodbox": "foodbox", "plate": "plate"}[g["env"]]](render=False, steps=2, path=[0, 0, 0], max_forces=F, num_interp_f=7)
    env.reset()
    _, r1, d1, i1 = env.step(g["u"][0])
    _, r2, d2, i2 = env.step(g["u"][1])
    assert (d1, d2) == (False, True) and i2["metric"] == float(metric.item())
    assert abs(-(r1 + r2) - float(J.item())) <= 1e-9 * max(1.0, abs(float(J.item())))
    with pytest.raises(RuntimeError):
        env.step(g["u"][1])


@pytest.mark.parametrize("env,F", [("bookshelf", 50.0), ("plate", 20.0)])
def test_phased_sorted_rollouts_equal_persistent_kernel(EN, env, F):
    """pg_set_rollout_mode: the phased path (one launch per episode phase, rollouts re-sorted by solver work in between, state in
    global memory) and the persistent one-launch kernel run the same per-rollout code -- poses, fingertip points, metric and
    costs must be bit-identical, also for a batch that leaves a partly filled CTA and for several problems in one launch."""
    from synthesize_pregrasp_b200.envspec import make_spec
    eng = EN.RolloutEngine(make_spec(env), device=0)
    K, N = 600, 2
    gen = torch.Generator(device="cuda:0").manual_seed(4242)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = 0.1 * np.random.RandomState(3).randn(N, 48)
    res = {}
    for mode in (1, 2):
        eng.set_rollout_mode(mode)
        out = eng.physics(u, noise, [0, 0], F)
        J, metric = eng.rollout(u, noise, [0, 0], F, want_metric=True)
        res[mode] = [out["pose"].cpu().numpy(), out["fins"].cpu().numpy(), out["metric"].cpu().numpy(), J.cpu().numpy(), metric.cpu().numpy()]
    for a, b in zip(res[1], res[2]):
        assert np.array_equal(a, b, equal_nan=True)
    # several MPPI problems in one launch (pg_rollout_multi) through the phased path
    Q = 3
    uQ = torch.as_tensor(0.1 * np.random.RandomState(4).randn(Q, N, 48), device="cuda:0")
    prob = torch.arange(Q, dtype=torch.int32, device="cuda:0").repeat_interleave(K // Q).contiguous()
    paths = np.zeros((Q, N), np.int32)
    Jm = {}
    for mode in (1, 2):
        eng.set_rollout_mode(mode)
        Jm[mode] = eng.rollout_multi(uQ, noise, prob, paths, F).cpu().numpy()
    assert np.array_equal(Jm[1], Jm[2], equal_nan=True)
    eng.set_rollout_mode(0)


def test_mppi_update_kernel(EN):
    from oracle import mppi
    m = np.load(os.path.join(G, "mppi_update.npz"))
    E = np.ascontiguousarray(np.transpose(m["E"], (2, 0, 1)), np.float32)
    u_t = torch.as_tensor(m["u0"], device="cuda:0").clone()
    EN.mppi_update(torch.as_tensor(m["J"], device="cuda:0"), torch.as_tensor(E, device="cuda:0"), u_t, 5.0, 1.0)
    ref, _ = mppi.mppi_update(m["u0"], m["J"], E, 5.0, 1.0)
    np.testing.assert_allclose(u_t.cpu().numpy(), ref, rtol=0, atol=1e-13)
    # edge cases: K=1, +inf costs, large K with a size-independent property (uniform J -> mean of the noise)
    rng = np.random.RandomState(0)
    for K in (1, 33, 100003):
        J = rng.normal(0, 1, K); E = rng.randn(K, 2, 48).astype(np.float32)
        if K > 1:
            J[::7] = np.inf
        u_t = torch.zeros((2, 48), dtype=torch.float64, device="cuda:0")
        EN.mppi_update(torch.as_tensor(J, device="cuda:0"), torch.as_tensor(E, device="cuda:0"), u_t, 5.0, 1.0)
        ref, _ = mppi.mppi_update(np.zeros((2, 48)), J, E, 5.0, 1.0)
        np.testing.assert_allclose(u_t.cpu().numpy(), ref, rtol=0, atol=1e-12)
    # every horizon the rollout ABI accepts (N <= 32): N*D = 48 ... 1536 columns, vectorised and scalar column paths
    for N, D in ((1, 48), (3, 48), (8, 48), (32, 48), (2, 7), (1, 1)):
        K = 777
        J = rng.normal(0, 1, K); E = rng.randn(K, N, D).astype(np.float32)
        J[5::11] = np.inf
        u0 = rng.randn(N, D)
        u_t = torch.as_tensor(u0, device="cuda:0").clone()
        EN.mppi_update(torch.as_tensor(J, device="cuda:0"), torch.as_tensor(E, device="cuda:0"), u_t, 5.0, 1.0)
        ref, _ = mppi.mppi_update(u0, J, E, 5.0, 1.0)
        np.testing.assert_allclose(u_t.cpu().numpy(), ref, rtol=0, atol=1e-12, err_msg=f"N={N} D={D}")
    K = 1 << 20
    E = torch.randn((K, 2, 48), device="cuda:0", dtype=torch.float32)
    u_t = torch.zeros((2, 48), dtype=torch.float64, device="cuda:0")
    EN.mppi_update(torch.full((K,), 3.0, dtype=torch.float64, device="cuda:0"), E, u_t, 5.0, 1.0)
    np.testing.assert_allclose(u_t.cpu().numpy(), E.double().mean(0).cpu().numpy(), rtol=0, atol=1e-12)


def test_dyn_feasibility_kernels(EN):
    from oracle import dynfeas
    from synthesize_pregrasp_b200 import dyn_feasibility_check as DF
    d = np.load(os.path.join(G, "dyn_simple.npz"))
    got = DF.simple_check_dyn_feasible(d["pts"], d["nrm"])
    bad = np.nonzero(got != d["out"])[0]
    # decisions are comparisons of fp64 expressions; they must agree with the reference function wherever the
    # decision is well conditioned, i.e. does not flip under a 1e-13 perturbation of the inputs in the oracle itself
    rng0 = np.random.RandomState(0)
    for b in bad:
        flips = {dynfeas.simple_check_dyn_feasible(d["pts"][b] + 1e-13 * rng0.randn(3, 3), d["nrm"][b] + 1e-13 * rng0.randn(3, 3)) for _ in range(16)}
        assert len(flips) == 2, f"item {b}: kernel {got[b]} reference {d['out'][b]} and the decision is well conditioned"
    print("dyn simple: mismatches on ill-conditioned items:", len(bad), "of", len(got))
    assert len(bad) <= 0.02 * len(got)
    gotm = DF.simple_check_dyn_feasible(d["m_pts"], d["m_nrm"], count=d["m_count"])
    assert (gotm != d["m_out"]).sum() <= 1
    assert DF.simple_check_dyn_feasible(d["pts"][0], d["nrm"][0]) == bool(d["out"][0])
    # sweep over a small cloud vs the oracle, full and split in two index ranges (multi-GPU partition)
    rng = np.random.RandomState(2)
    P = 40
    box = np.array([0.2, 0.2, 0.05])
    pts = np.zeros((P, 3)); nrm = np.zeros((P, 3))
    for i in range(P):
        ax = rng.randint(3); sg = rng.choice([-1.0, 1.0]); q = rng.uniform(-1, 1, 3) * box; q[ax] = sg * box[ax]
        pts[i] = q; nrm[i, ax] = -sg
    import itertools
    ref = [c for c in itertools.combinations(range(P), 3)
           if min(np.linalg.norm(pts[a] - pts[b]) for a, b in itertools.combinations(c, 2)) >= 0.03
           and dynfeas.simple_check_dyn_feasible_3p(pts[list(c)], nrm[list(c)])]
    tri, n = DF.find_dynamically_feasible_contacts(pts, nrm)
    assert n == len(ref) and sorted(map(tuple, tri.tolist())) == sorted(ref)
    total = P * (P - 1) * (P - 2) // 6
    t1, n1 = DF.find_dynamically_feasible_contacts(pts, nrm, first=0, last=total // 2)
    t2, n2 = DF.find_dynamically_feasible_contacts(pts, nrm, first=total // 2, last=total)
    assert n1 + n2 == n and sorted(map(tuple, np.concatenate([t1, t2]).tolist())) == sorted(ref)


def test_optimizer_api_runs_and_improves(EN):
    """StochTrajOptimizer with the reference's constructor/return contract (stoch_traj_opt.py:31-46, :231)."""
    from synthesize_pregrasp_b200.envs import envs_dict
    from synthesize_pregrasp_b200.stoch_traj_opt import StochTrajOptimizer
    opt = StochTrajOptimizer(env=envs_dict["foodbox"], sigma=0.8, initial_guess=None, TimeSteps=2, seed=12367134,
                             render=False, Iterations=3, active_finger_tips=[0, 1], num_interp_f=7, Num_processes=12,
                             Traj_per_process=15, opt_time=False, verbose=0, mode="only_score", steps=2,
                             path=np.array([0, 0, 0]), has_distance_field=True, max_forces=20.0, validate=False)
    uopt, neg_jopt, jopt_log, r_log = opt.optimize()
    assert uopt.shape == (2, 48) and len(jopt_log) == 3 and len(r_log) == 3
    assert np.isfinite(neg_jopt) and jopt_log[-1] >= jopt_log[0]
    J, metric = opt.replay_traj(uopt)
    assert abs(-J - neg_jopt) <= 1e-9 * max(1.0, abs(J))


def test_model_optimize_driver_serial_and_batched(EN, tmp_path):
    """Host mirror of model_optimize.py:36-123 (paths x runs loops, output files), both schedules, tiny sizes."""
    from synthesize_pregrasp_b200 import model_optimize as MO
    base = ["--env", "foodbox", "--max_force", "20", "--has_distance_field", "--iters", "2", "--runs", "2", "--exp_name", "t"]
    res = MO.run(MO.build_parser().parse_args(base + ["--out_dir", str(tmp_path / "serial")]))
    assert len(res) == 1 and res[0]["uopt"].shape == (2, 48) and len(res[0]["Jopt_log"]) == 2
    assert np.isfinite(res[0]["best"]) and res[0]["best"] >= res[0]["mean"] - 1e-9
    for f in ("rewards/optimal_t_20.0.npy", "rewards/run_t_20.0.npy", "traj/t_20.0.npy", "log/t.txt"):
        assert (tmp_path / "serial" / f).exists(), f
    np.testing.assert_array_equal(np.load(tmp_path / "serial" / "traj" / "t_20.0.npy"), res[0]["uopt"])
    # batched schedule: run j of the path is seeded like the serial run j, and every problem of a multi-problem launch
    # reproduces its serial optimiser -> same best reward and controls
    resb = MO.run(MO.build_parser().parse_args(base + ["--batched", "--out_dir", str(tmp_path / "batched")]))
    assert len(resb) == 1 and abs(resb[0]["best"] - res[0]["best"]) <= 1e-9 * max(1.0, abs(res[0]["best"]))
    np.testing.assert_allclose(resb[0]["uopt"], res[0]["uopt"], rtol=0, atol=1e-12)
    # --validate takes its paths from paths_id.npy and the committed contact-state graph
    resv = MO.run(MO.build_parser().parse_args(["--env", "plate", "--max_force", "20", "--has_distance_field", "--validate", "--iters", "1",
                                                "--runs", "1", "--max_paths", "2", "--batched", "--exp_name", "v",
                                                "--out_dir", str(tmp_path / "validate")]))
    assert [r["path"] for r in resv] == [[1, 2], [1, 7]] and all(np.isfinite(r["best"]) for r in resv)
    # every path keeps its own controls and model_test recording (the un-suffixed files hold the path optimised last, as the
    # reference's would when no path passes its validate_path screening, model_optimize.py:97-117)
    for i, r in enumerate(resv):
        np.testing.assert_array_equal(np.load(tmp_path / "validate" / "traj" / f"v_path{i}_20.0.npy"), r["uopt"])
        assert np.load(tmp_path / "validate" / "object_poses" / f"v_path{i}_20.0_object_poses.npy").shape == (100, 7)
    np.testing.assert_array_equal(np.load(tmp_path / "validate" / "traj" / "v_20.0.npy"), resv[-1]["uopt"])


def test_model_test_recording_mirror(EN, tmp_path):
    """model_test.py:115-135 without --add_physics: one replay, [N*50,7] object poses and [N*50,5,6] fingertip records."""
    from synthesize_pregrasp_b200 import model_test as MT
    g = recorded("foodbox_20.0")
    raw = np.load(os.path.join(ROOT, "tests", "golden", "recorded_tip_records.npz"))["foodbox_20p0"]
    os.makedirs(tmp_path / "traj")
    np.save(tmp_path / "traj" / "foodbox_20.0.npy", g["u"])
    MT.main(["--exp_name", "foodbox_20.0", "--env", "foodbox", "--data_dir", str(tmp_path)])
    poses = np.load(tmp_path / "object_poses" / "foodbox_20.0_object_poses.npy")
    tips = np.load(tmp_path / "tip_data" / "foodbox_20.0_tip_poses.npy")
    assert poses.shape == (100, 7) and tips.shape == (100, 5, 6) and np.all(tips[:, 2:] == 100.0)
    # the first substeps reproduce the PyBullet recording (the later ones as far as the oracle does)
    assert np.abs(poses[:6] - g["poses"][:6]).max() <= 1e-7
    assert np.abs(tips[:6, :2] - raw[:6, :2]).max() <= 1e-6
    # every row: the record is the object pose applied to the decoded local contact point
    tips2, poses2, J, metric = MT.record("foodbox", g["u"], 20.0)
    np.testing.assert_array_equal(poses2, poses)
    np.testing.assert_array_equal(tips2, tips)
    assert np.isfinite(J) and np.isfinite(metric)


def test_errors_are_reported_not_swallowed(EN):
    from synthesize_pregrasp_b200.envspec import make_spec
    import dataclasses
    with pytest.raises(RuntimeError):
        EN.RolloutEngine(dataclasses.replace(make_spec("plate"), obj_hull=None), device=0)    # hull object without vertices
    eng = EN.RolloutEngine(make_spec("foodbox"), device=0)
    with pytest.raises(RuntimeError):
        eng.rollout(np.zeros((2, 48)), None, [0, 7], 20.0, K=1)   # contact state out of range


def test_ragged_and_edge_batches(EN):
    """K not a multiple of the CTA size, K=1, N=1 and N=3: each rollout must equal the same rollout run alone."""
    from synthesize_pregrasp_b200.envspec import make_spec
    eng = EN.RolloutEngine(make_spec("keyboard"), device=0)
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    for K, N in ((1, 2), (65, 1), (130, 3)):
        noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
        u = 0.1 * np.ones((N, 48))
        J, metric = eng.rollout(u, noise, [0] * N, 20.0, want_metric=True)
        assert J.shape == (K,) and torch.isfinite(J).all()
        for k in (0, K - 1):
            Jk, mk = eng.rollout(u, noise[k:k + 1].contiguous(), [0] * N, 20.0, want_metric=True)
            assert float(Jk) == float(J[k]) and float(mk) == float(metric[k])      # batch position must not matter
    # replay form (noise = NULL) equals zero noise
    z = torch.zeros((1, 2, 48), device="cuda:0", dtype=torch.float32)
    u = 0.3 * np.ones((2, 48))
    assert float(eng.rollout(u, None, [0, 0], 20.0, K=1)) == float(eng.rollout(u, z, [0, 0], 20.0))


def test_two_env_handles_on_two_streams(EN):
    """BASELINE config 5: independent env handles run concurrently on their own streams."""
    from synthesize_pregrasp_b200.envspec import make_spec
    engs = [EN.RolloutEngine(make_spec(n), device=0) for n in ("foodbox", "keyboard")]
    gen = torch.Generator(device="cuda:0").manual_seed(8)
    noise = 0.8 * torch.randn((96, 2, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    u = np.zeros((2, 48))
    ref = [e.rollout(u, noise, [0, 0], 20.0).clone() for e in engs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in engs]
    outs = []
    for e, s in zip(engs, streams):
        with torch.cuda.stream(s):
            outs.append(e.rollout(u, noise, [0, 0], 20.0))
    torch.cuda.synchronize()
    for a, b in zip(ref, outs):
        assert torch.equal(a, b)


def test_two_devices_in_one_process(EN):
    """pg_env_create(spec, device): engines on two GPUs of one process (per-device kernel attributes, per-handle scratch, the MPPI
    update on the second device) give the same results bit for bit.  Needs two visible GPUs (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from synthesize_pregrasp_b200.envspec import make_spec
    outs = []
    for dev in (1, 0):                                            # the second device first: nothing may rely on device 0's state
        eng = EN.RolloutEngine(make_spec("plate"), device=dev)
        gen = torch.Generator(device=f"cuda:{dev}").manual_seed(21)
        noise = 0.8 * torch.randn((200, 2, 48), generator=gen, device=f"cuda:{dev}", dtype=torch.float32)
        u = torch.zeros((2, 48), dtype=torch.float64, device=f"cuda:{dev}")
        J = eng.rollout(u, noise, [0, 0], 20.0)
        with torch.cuda.device(dev):
            EN.mppi_update(J, noise, u, 5.0, 1.0)
        torch.cuda.synchronize(dev)
        outs.append((noise.cpu(), J.cpu(), u.cpu()))
        Jh = eng.rollout_host(np.zeros((2, 48)), noise.cpu().numpy(), [0, 0], 20.0)
        assert np.array_equal(Jh, J.cpu().numpy())
        eng.close()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_csg_node_scores_match_oracle(EN):
    """generate_csg.node_scores: all (node x sample) score evaluations in one pg_score launch vs the oracle network."""
    from oracle import scorenet as SN
    from synthesize_pregrasp_b200 import generate_csg as G
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec("plate")
    eng = EN.RolloutEngine(spec, device=0)
    rng = np.random.RandomState(11)
    cloud = spec.cloud[:1024].astype(np.float32)
    df = (cloud[:, 2] + 0.1).astype(np.float32)                 # generate_csg.py:47
    regions = list(range(6))
    samples = [cloud[rng.randint(0, 1024, size=G.NUM_SAMPLES)] + 0.002 * rng.randn(G.NUM_SAMPLES, 3).astype(np.float32) for _ in regions]
    nodes = G.all_ordered_nodes(regions)
    got = G.node_scores(eng, cloud, df, samples, nodes)
    w = oracle_weights()
    for n in (0, 7, len(nodes) - 1):
        i, j = nodes[n]
        ref = np.mean([SN.score_logit(w, cloud.astype(np.float64), df.astype(np.float64), np.r_[samples[i][t], samples[j][t]].astype(np.float64))
                       for t in range(G.NUM_SAMPLES)])
        assert abs(got[n] - ref) <= 2e-5 * max(1.0, abs(ref)), (n, got[n], ref)
    table, ids, _, _ = G.select_paths(nodes[:10], got[:10], nodes, got)
    assert ids.shape == (10, 2) and table.shape[1] == 2


def test_multi_problem_rollouts_equal_separate_problems(EN):
    """pg_rollout_multi: Q problems (own controls, own path) in one launch == Q single-problem launches, bit for bit; the
    multi-problem optimiser reproduces StochTrajOptimizer per problem (model_optimize.py:74-117 runs them serially)."""
    from synthesize_pregrasp_b200.envs import envs_dict
    from synthesize_pregrasp_b200.envspec import make_spec
    from synthesize_pregrasp_b200.stoch_traj_opt import MultiStochTrajOptimizer, StochTrajOptimizer
    spec = make_spec("plate")
    eng = EN.RolloutEngine(spec, device=0)
    gen = torch.Generator(device="cuda:0").manual_seed(9)
    Q, M, N = 3, 40, 2
    u = 0.3 * torch.randn((Q, N, 48), generator=gen, device="cuda:0", dtype=torch.float64)
    noise = 0.8 * torch.randn((Q * M, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    n_states = len(spec.dummy_states)
    paths = np.array([[q % n_states, (q + 1) % n_states] for q in range(Q)], np.int32)
    prob = torch.arange(Q, dtype=torch.int32, device="cuda:0").repeat_interleave(M).contiguous()
    J, metric = eng.rollout_multi(u, noise, prob, paths, 20.0, want_metric=True)
    for q in range(Q):
        Jq, mq = eng.rollout(u[q], noise[q * M:(q + 1) * M].contiguous(), paths[q], 20.0, want_metric=True)
        assert torch.equal(J[q * M:(q + 1) * M], Jq) and torch.equal(metric[q * M:(q + 1) * M], mq), q
    # interleaved problem indices: same rollouts, permuted
    perm = torch.randperm(Q * M, generator=torch.Generator().manual_seed(1)).to("cuda:0")
    J2 = eng.rollout_multi(u, noise[perm].contiguous(), prob[perm].contiguous(), paths, 20.0)
    assert torch.equal(J2, J[perm])
    kw = dict(sigma=0.8, TimeSteps=2, Iterations=2, active_finger_tips=[0, 1], num_interp_f=7, Num_processes=4, Traj_per_process=10,
              opt_time=False, mode="only_score", steps=2, has_distance_field=True, max_forces=20.0, validate=False)
    seeds = [11, 12]
    multi = MultiStochTrajOptimizer(envs_dict["foodbox"], paths=[[0, 0, 0], [0, 0, 0]], seeds=seeds, path=np.array([0, 0, 0]), **kw)
    uopt, ropt, logs = multi.optimize()
    for q, sd in enumerate(seeds):
        single = StochTrajOptimizer(env=envs_dict["foodbox"], seed=sd, render=False, initial_guess=None, verbose=0,
                                    path=np.array([0, 0, 0]), **kw)
        u1, r1, log1, _ = single.optimize()
        np.testing.assert_allclose(uopt[q], u1, rtol=0, atol=1e-12)
        assert abs(ropt[q] - r1) <= 1e-9 * max(1.0, abs(r1))


@pytest.mark.parametrize("env,F", [("plate", 20.0), ("bookshelf", 50.0)])
def test_settle_regrouping_does_not_change_results(EN, env, F):
    """Large batches re-deal the rollouts of a CTA to its threads before the settle phase (csrc/pg_rollout.cuh
    settle_regroup); batches that do not fill the lanes of 14 warps per SM do not.  A rollout's result must not depend on it: one launch of K
    rollouts == the same rollouts in small launches, bit for bit (cost J goes through pose and fingertip outputs)."""
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(env)
    eng = EN.RolloutEngine(spec, device=0)
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    K, N = n_sm * 448 - 7, 2                        # full lanes, 448-thread CTAs, the last one partial (no regrouping there)
    u = 0.2 * torch.randn((N, 48), generator=gen, device="cuda:0", dtype=torch.float64)
    noise = 0.8 * torch.randn((K, N, 48), generator=gen, device="cuda:0", dtype=torch.float32)
    path = np.zeros(N, np.int32)
    J, metric = eng.rollout(u, noise, path, F, want_metric=True)
    chunk = 8192
    for a in range(0, K, chunk):
        Jc, mc = eng.rollout(u, noise[a:a + chunk].contiguous(), path, F, want_metric=True)
        assert torch.equal(J[a:a + chunk], Jc) and torch.equal(metric[a:a + chunk], mc), a
    assert torch.isfinite(J).all() and J.std() > 0


def test_dyn_feasibility_qp_kernels(EN):
    """pg_dyn_feasible_qp / pg_dyn_sweep_qp vs the oracle's independent solve of the reference QP (parity unpinned vs cvxpy)."""
    from oracle import dynfeas as DF
    from synthesize_pregrasp_b200 import dyn_feasibility_check as DYN
    rng = np.random.RandomState(21)
    B, M = 60, 4
    pts = rng.uniform(-0.2, 0.2, (B, M, 3))
    nrm = rng.randn(B, M, 3)
    nrm[::2] = pts[::2] - pts[::2].mean(1, keepdims=True)        # closure-like half
    pts[5, 3, 0] = 100.0                                          # an absent finger (x >= 50)
    dec, obj = DYN.check_dyn_feasible(pts, nrm, tol=2e-3, return_objective=True)
    n_close = 0
    for b in range(B):
        rd, ro = DF.check_dyn_feasible_qp(pts[b], nrm[b], tol=2e-3)
        if abs(ro - 2e-3) <= 1e-5:                                # within solver noise of the threshold: either decision is defensible
            n_close += 1
            continue
        assert bool(dec[b]) == rd, (b, obj[b], ro)
        assert abs(obj[b] - ro) <= 1e-6 * max(1.0, ro) + 1e-8, (b, obj[b], ro)
    assert n_close <= 3 and 5 < dec.sum() < B
    one = DYN.check_dyn_feasible(pts[0], nrm[0], tol=2e-3)
    assert one == bool(dec[0])
    # sweep with the QP criterion == brute force over all triples of a small cloud, also when split in two ranges
    P = 14
    cp = rng.uniform(-0.15, 0.15, (P, 3))
    cn = cp - cp.mean(0) + 0.05 * rng.randn(P, 3)
    cn /= np.linalg.norm(cn, axis=1)[:, None]
    tri, n = DYN.find_dynamically_feasible_contacts(cp, cn, criterion="qp", tol=1e-4)
    import itertools
    want = set()
    for c in itertools.combinations(range(P), 3):
        q = cp[list(c)]
        if min(np.linalg.norm(q[0] - q[1]), np.linalg.norm(q[0] - q[2]), np.linalg.norm(q[1] - q[2])) < 0.03:
            continue
        o = DF.qp_objective(q, cn[list(c)], normalise=False)
        if abs(o - 1e-4) <= 1e-6:
            continue
        if o <= 1e-4:
            want.add(c)
    got = {tuple(t) for t in tri.tolist()}
    assert want <= got and len(got - want) <= 2, (len(want), len(got))
    total = P * (P - 1) * (P - 2) // 6
    t1, n1 = DYN.find_dynamically_feasible_contacts(cp, cn, criterion="qp", tol=1e-4, first=0, last=total // 2)
    t2, n2 = DYN.find_dynamically_feasible_contacts(cp, cn, criterion="qp", tol=1e-4, first=total // 2)
    assert n1 + n2 == n and {tuple(t) for t in t1.tolist()} | {tuple(t) for t in t2.tolist()} == got
