"""bench.py -- rollout-steps/s of the MPPI rollout hot path on N GPUs of one node (see the run contract in
DESIGN.md "Measurement").  One step = one pass of the hot path over one batch of K synthetic rollouts:
K x (reset + 2 env steps of 1+50 substeps + settle + 2 cost evaluations).
  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
  python bench.py --impl reference ...   (the CPU restatement of the reference path on the box's host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {   # BASELINE.json configs: env, max force, K per GPU
    "bookshelf_K65536": dict(env="bookshelf", force=50.0, K=65536),
    "foodbox_K65536": dict(env="foodbox", force=20.0, K=65536),
    "plate_K65536": dict(env="plate", force=20.0, K=65536),        # configs[0]'s task at the headline batch size
    "plate_K360": dict(env="plate", force=20.0, K=360),            # configs[0] as the reference runs it: 12 workers x 30 (model_optimize.py:77)
    "waterbottle_K65536": dict(env="waterbottle", force=20.0, K=65536),   # SURVEY 8(d) config 4, per-GPU share
    # the reference's whole plate job per MPPI iteration: 10 contact-state paths x 3 runs (model_optimize.py:74-117), each with
    # its 12 x 30 samples, as ONE multi-problem launch (pg_rollout_multi) instead of 30 serial optimisers
    "plate_Q30xK360": dict(env="plate", force=20.0, K=30 * 360, Q=30),
    # SURVEY 8(d) config 4 in its strong-scaling form: ONE job of 2^20 samples split over the ranks, one all_gather per iteration
    "waterbottle_K1048576": dict(env="waterbottle", force=20.0, K=1 << 20, strong=True),
    "bookshelf_K1048576": dict(env="bookshelf", force=50.0, K=1 << 20, strong=True),
}
SIGMA, KAPPA, ALPHA, N_STEPS, D = 0.8, 5.0, 1.0, 2, 48


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, src="fallback")


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons}


_POOL_STATE = {}


def _pool_init(env, path=None, validate=False):
    try:                                   # one BLAS thread per worker process: the workers are the parallelism
        from threadpoolctl import threadpool_limits
        _POOL_STATE["blas_limit"] = threadpool_limits(1)
    except Exception:
        pass
    from helpers import oracle_cost_spec, oracle_spec, oracle_weights
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(env)
    states = spec.csg_states if validate else None
    p = list(path) if path is not None else [0] * N_STEPS
    _POOL_STATE.update(osd=oracle_spec(spec, states), ocs=oracle_cost_spec(spec), w=oracle_weights(), path=p + [p[-1]])


def _pool_one(args):
    from oracle.rollout import rollout as oracle_rollout
    force, nz = args
    st = _POOL_STATE
    # fp32 network arithmetic, as the reference's torch module computes it
    return oracle_rollout(st["osd"], st["ocs"], st["w"], np.zeros((N_STEPS, D)), nz, st["path"], force, net_dtype=np.float32)["J"]


class CpuReference:
    """The CPU restatement (oracle/) of the same path on host cores -- the reported baseline / reference arm.
    This is the one place bench.py executes oracle/, never as the product.  A persistent process pool is
    created and warmed once so that the timed region holds rollouts only."""

    def __init__(self, env, force, threads, path=None, validate=False):
        import concurrent.futures as cf
        self.env, self.force, self.threads = env, force, threads
        self.pool = cf.ProcessPoolExecutor(threads, initializer=_pool_init, initargs=(env, path, validate)) if threads > 1 else None
        if self.pool is None:
            _pool_init(env, path, validate)
        self.run(max(threads, 1), seed=12345)          # warm every worker (imports, library load)

    def run(self, n_rollouts, seed=0):
        rng = np.random.RandomState(seed)
        noise = (SIGMA * rng.randn(n_rollouts, N_STEPS, D)).astype(np.float32)
        t0 = time.time()
        if self.pool is None:
            for k in range(n_rollouts):
                _pool_one((self.force, noise[k]))
        else:
            list(self.pool.map(_pool_one, [(self.force, noise[k]) for k in range(n_rollouts)], chunksize=max(1, n_rollouts // (4 * self.threads))))
        dt = time.time() - t0
        return n_rollouts * N_STEPS / dt, dt

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()


# --------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 5: not rollout-steps of one env, so they print their own metric on the same contract
def _box_cloud(P, seed=0, size=(0.4, 0.4, 0.1)):
    """Surface cloud of the 0.4 x 0.4 x 0.1 box of utils/dyn_feasibility_check.py:312-317 (area-weighted uniform sample)."""
    rng = np.random.RandomState(seed)
    h = np.array(size) / 2
    area = np.array([size[1] * size[2], size[0] * size[2], size[0] * size[1]])
    ax = rng.choice(3, size=P, p=area / area.sum()); sgn = rng.choice([-1.0, 1.0], size=P)
    pts = rng.uniform(-h, h, size=(P, 3)); nrm = np.zeros((P, 3))
    pts[np.arange(P), ax] = sgn * h[ax]; nrm[np.arange(P), ax] = sgn
    return pts, nrm


def bench_dyn_sweep(args, rank, world, local):
    """Config 3-i: find_dynamically_feasible_contacts over ALL C(P,3) triples of a P-point cloud, simple criterion and the
    wrench-closure QP (the criterion the reference runs, one cvxpy solve per triple), triple index space split over the ranks."""
    import torch
    import torch.distributed as dist
    from synthesize_pregrasp_b200 import _abi
    from synthesize_pregrasp_b200 import dyn_feasibility_check as DYN
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    P = int(args.workload.split("_P")[1])
    pts, nrm = _box_cloud(P)
    total = P * (P - 1) * (P - 2) // 6
    first, last = DYN.triple_range(rank, world, P)
    L = _abi.lib()
    res = {}
    for crit in ("qp", "simple"):
        for _ in range(args.warmup):
            DYN.find_dynamically_feasible_contacts(pts, nrm, first=first, last=last, criterion=crit, max_out=1 << 20, tol=1e-4)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = L.pg_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            tri, n = DYN.find_dynamically_feasible_contacts(pts, nrm, first=first, last=last, criterion=crit, max_out=1 << 20, tol=1e-4)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps, float(n)], dtype=torch.float64, device="cuda")
        if world > 1:
            tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            t = torch.stack([tm[0], ts[1]])
        res[crit] = dict(ms=float(t[0]), feasible=int(t[1]), launches=int(L.pg_launch_count() - l0))
    if rank == 0:
        # CPU baseline: the oracle's QP (scipy SLSQP on the reference's own Q, G, h) and simple test on a bounded sample of triples
        from oracle import dynfeas
        rng = np.random.RandomState(0)
        idx = np.array([sorted(rng.choice(P, 3, replace=False)) for _ in range(200)])
        t0 = time.perf_counter()
        for i, j, k in idx:
            dynfeas.qp_objective(pts[[i, j, k]], nrm[[i, j, k]])
        cpu_qp = len(idx) / (time.perf_counter() - t0)
        f64_peak, f64_src = _fp64_peak()
        QP_FLOP = 20e3                      # ~10 Lawson-Hanson outer iterations of a 9 x 15 NNLS (DESIGN.md kernel 5)
        v = total / (res["qp"]["ms"] * 1e-3)
        ach = v * QP_FLOP / 1e12
        print(json.dumps({"metric": "feasibility-triples/sec (wrench-closure QP)", "value": v, "unit": "triples/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": res["qp"]["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": args.workload, "points": P, "triples": total, "tol": 1e-4, "min_dist": 0.03,
                                     "cloud": "0.4 x 0.4 x 0.1 box surface, seed 0", "partition": "contiguous triple-index ranges, one per rank"},
                          "feasible_qp": res["qp"]["feasible"], "feasible_simple": res["simple"]["feasible"],
                          "simple_criterion": {"ms_per_step": res["simple"]["ms"], "triples_per_s": total / (res["simple"]["ms"] * 1e-3)},
                          "roofline": {"kernel": "dyn_sweep_qp_kernel", "bound": "fp64-issue", "achieved": ach, "peak": f64_peak, "unit": "TFLOP/s",
                                       "frac": ach / f64_peak, "traffic": None, "peak_source": f64_src,
                                       "note": "~20 kFLOP per triple (estimate); triples are enumerated on the fly, HBM sees only the cloud and the hits"},
                          "cpu_baseline": {"value": cpu_qp, "unit": "triples/s", "cores": 1, "kind": "port",
                                           "sample": "200 random triples through oracle/dynfeas.qp_objective (scipy SLSQP on the reference's Q, G, h; the reference itself pays a cvxpy solve per triple)"},
                          "e2e": {"value": v, "unit": "triples/s", "h2d_bytes_per_step": int(pts.nbytes + nrm.nbytes), "d2h_bytes_per_step": int(res["qp"]["feasible"] * 12 + 8)},
                          "gpu_launches": res["qp"]["launches"], "host_cores": os.cpu_count() or 1}))
    if world > 1:
        dist.destroy_process_group()


MULTI_ENVS = (("plate", 20.0), ("bookshelf", 50.0), ("laptop", 20.0), ("keyboard", 20.0), ("cardboard", 20.0))


def bench_multi_env(args, rank, world, local):
    """Config 5: plate, bookshelf, laptop, keyboard, cardboard optimised CONCURRENTLY -- one handle and one CUDA stream per env on
    every rank -- over a sweep of sample counts K per env.  One JSON line per K (rank 0); value = rollout-steps/s of all envs."""
    import torch
    import torch.distributed as dist
    from synthesize_pregrasp_b200 import _abi
    from synthesize_pregrasp_b200 import engine as EN
    from synthesize_pregrasp_b200.envspec import make_spec
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    L = _abi.lib()
    engines, streams, paths = [], [], []
    for env, force in MULTI_ENVS:
        spec = make_spec(env)
        path, validate = _bench_path(spec)
        engines.append(EN.RolloutEngine(spec, device=local, validate=validate)); paths.append(path)
        streams.append(torch.cuda.Stream(device=dev))
    sweep = [args.K] if args.K else [1024, 4096, 16384, 65536]
    u = torch.zeros((N_STEPS, D), dtype=torch.float64, device=dev)
    for K in sweep:
        gen = torch.Generator(device=dev).manual_seed(12367134 + rank)
        noise = [SIGMA * torch.randn((K, N_STEPS, D), generator=gen, device=dev, dtype=torch.float32) for _ in MULTI_ENVS]

        def step():
            for i, (env, force) in enumerate(MULTI_ENVS):
                with torch.cuda.stream(streams[i]):
                    J = engines[i].rollout(u, noise[i], paths[i], force)
                    EN.mppi_update(J, noise[i], None, KAPPA, ALPHA, apply=False, want_partial=True)
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        l0 = L.pg_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            for st in streams:
                st.wait_stream(torch.cuda.current_stream())
            step()
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = int(L.pg_launch_count() - l0)
        if rank == 0:
            ms = float(t[0]); units = K * world * N_STEPS * len(MULTI_ENVS)
            print(json.dumps({"metric": "rollout-steps/sec (five envs concurrently)", "value": units / (ms * 1e-3), "unit": "rollout-steps/s", "n_gpus": world,
                              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                              "config": {"workload": "multi_env", "envs": [e for e, _ in MULTI_ENVS], "K_per_env_per_gpu": K, "N": N_STEPS, "D": D,
                                         "streams": len(MULTI_ENVS), "l2": "flushed by a 256 MB write before every timed step"},
                              "roofline": None, "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "host_cores": os.cpu_count() or 1}))
    if world > 1:
        dist.destroy_process_group()


def bench_mppi_update(args, rank, world, local):
    """Row a2 alone: the exp-weighted update (stoch_traj_opt.py:195-204) over K = 2^20 samples per GPU -- the one kernel of the path
    that is bound by HBM (it streams the noise once): min J, exp weights, sum of w * noise[K,N,D], update of u.  Every rank runs its
    own K (the multi-GPU merge is 99 doubles per rank and is part of the rollout workloads)."""
    import torch
    import torch.distributed as dist
    from synthesize_pregrasp_b200 import _abi
    from synthesize_pregrasp_b200 import engine as EN
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = args.K or (1 << 20)
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    noise = SIGMA * torch.randn((K, N_STEPS, D), generator=gen, device=dev, dtype=torch.float32)      # 403 MB > 126 MB L2
    J = 1000.0 + 10.0 * torch.rand((K,), generator=gen, device=dev, dtype=torch.float64)
    u = torch.zeros((N_STEPS, D), dtype=torch.float64, device=dev)
    L = _abi.lib()
    for _ in range(max(3, args.warmup)):
        EN.mppi_update(J, noise, u, KAPPA, ALPHA)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    steps = max(args.steps, 5000)                       # ~0.1 ms each: long enough for the clock sampler to see the load
    l0 = L.pg_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        EN.mppi_update(J, noise, u, KAPPA, ALPHA)
    e1.record(); torch.cuda.synchronize()
    launches = int(L.pg_launch_count() - l0)
    sampler.stop_flag = True; sampler.join()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier(); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    # end to end: host J and pinned host noise in, updated u out
    noise_h = noise.cpu().pin_memory(); J_h = J.cpu().pin_memory()
    t0 = time.perf_counter()
    for _ in range(3):
        nd = noise_h.to(dev, non_blocking=True); Jd = J_h.to(dev, non_blocking=True)
        EN.mppi_update(Jd, nd, u, KAPPA, ALPHA); u_h = u.cpu()
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    if rank == 0:
        from oracle import mppi as OM
        ks = 1 << 16
        Jn, nn = J[:ks].cpu().numpy(), noise[:ks].cpu().numpy()
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 3.0:
            OM.mppi_update(np.zeros((N_STEPS, D)), Jn, nn, KAPPA, ALPHA); reps += 1
        cpu = reps * ks / (time.perf_counter() - t0)
        peaks = _peaks()
        bytes_alg = K * (N_STEPS * D * 4 + 2 * 8)               # the noise once, J twice (min pass, weight pass)
        ach = bytes_alg / (ms * 1e-3) / 1e9
        print(json.dumps({"metric": "mppi-update samples/sec", "value": K * world / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": steps,
                          "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64 accumulation over f32 noise", "data": "synthetic",
                          "config": {"workload": args.workload, "K_per_gpu": K, "N": N_STEPS, "D": D, "kappa": KAPPA,
                                     "l2": f"noise {K * N_STEPS * D * 4 / 1e6:.0f} MB > 126 MB L2"},
                          "roofline": {"kernel": "mppi_min + mppi_accum<float4> + mppi_final", "bound": "hbm", "achieved": ach, "peak": peaks["hbm"],
                                       "unit": "GB/s", "frac": ach / peaks["hbm"], "traffic": None, "peak_source": peaks["src"],
                                       "note": "algorithmic bytes = 4*N*D per sample (noise, read once) + 16 (J read by the min pass and the weight pass)"},
                          "cpu_baseline": {"value": cpu, "unit": "samples/s", "cores": 1, "kind": "port",
                                           "sample": f"{ks} samples through oracle/mppi.mppi_update (numpy: the reference's own lines) for 3 s"},
                          "e2e": {"value": K * world / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(noise_h.nbytes + J_h.nbytes),
                                  "d2h_bytes_per_step": int(u_h.numel() * 8)},
                          "gpu_launches": launches, "clocks": sampler.summary(),
                          "host_cores": os.cpu_count() or 1}))
    if world > 1:
        dist.destroy_process_group()


OTHER_WORKLOADS = {"mppi_update_K1048576": bench_mppi_update, "dyn_sweep_P256": bench_dyn_sweep, "dyn_sweep_P1024": bench_dyn_sweep, "multi_env": bench_multi_env}


def _lib_sha16():
    import hashlib
    from synthesize_pregrasp_b200 import _abi
    return hashlib.sha256(open(_abi.LIB_PATH, "rb").read()).hexdigest()[:16]


def _src_sha16():
    """sha256 over the sources the library is built from (csrc/*.cu, *.cuh, *.h, Makefile): what identifies 'this build' across
    machines -- nvcc's output is not byte-reproducible, the sources are."""
    import hashlib
    d = os.path.join(ROOT, "synthesize_pregrasp_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")) or f == "Makefile":
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "pregrasp_b200.h"), "rb").read())
    return h.hexdigest()[:16]


BOX_ROLLOUT_KERNEL = "_ZN2pg14rollout_kernelILi0EEEvPKNS_6EnvDevEPKdPKfiiPKidPdSA_SA_SA_PiS9_Pvi"      # pg::rollout_kernel<SHAPE_BOX>


def _kernel_sass_sha16(mangled):
    """sha256 over the SASS instruction lines of ONE kernel of the library being timed (cuobjdump -sass -fun): identifies the
    machine code of that kernel -- an edit to a header that only another kernel compiles (pg_convex.cuh: hull objects) changes the
    source hash but not this one.  None when cuobjdump is not on the box."""
    import hashlib
    import re
    import shutil
    import subprocess
    from synthesize_pregrasp_b200 import _abi
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", "-fun", mangled, _abi.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.SubprocessError):
        return None
    lines = [ln for ln in out.splitlines() if re.match(r"^\s+/\*[0-9a-f]{4}\*/", ln)]
    if not lines:
        return None
    return hashlib.sha256(("\n".join(lines) + "\n").encode()).hexdigest()[:16]


def _fp64_peak():
    """Measured DFMA rate of the box (tools/ubench -> profiles/r2a_ubench.json), else the nominal 148 SM x 64 DFMA/clk."""
    p = os.path.join(ROOT, "profiles", "r2a_ubench.json")
    if os.path.exists(p):
        return float(json.load(open(p))["dfma_peak_tflops"]), "measured (tools/ubench/ubench.cu, profiles/r2a_ubench.json)"
    return 148 * 64 * 2 * 1.965e9 / 1e12, "nominal 148 SM x 64 DFMA/clk x 1.965 GHz"


def _ncu_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE rollout_kernel launch of THIS build (profiles/ncu_traffic.json is
    written by tools/ncu_traffic.py from a capture of the same build, keyed by the sha256 of the library's sources); null otherwise."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    if d.get("src_sha16") != _src_sha16():
        # the sources changed since the capture: the figure still describes this build if the machine code of the captured kernel
        # is the same (recorded next to the capture as kernel_sass_sha16)
        want = d.get("kernel_sass_sha16", {}).get(workload)
        if want is None or _kernel_sass_sha16(d.get("kernel", {}).get(workload, BOX_ROLLOUT_KERNEL)) != want:
            return None
    return d.get("traffic_bytes", {}).get(workload)


def _bench_path(spec):
    """the first row of paths_id.npy under the committed csg.npy (model_optimize.py:65-66, --validate)."""
    if spec.paths_id is not None and spec.csg_states is not None:
        return [int(x) for x in spec.paths_id[0][:N_STEPS]], True
    return [0] * N_STEPS, False


def measure_rollouts(args, name, wl, steps, warmup, rank, world, local, cpu_seconds, cpu_sample=None):
    """One rollout workload on this rank's GPU -> dict of measurements (times are max over ranks)."""
    import torch
    import torch.distributed as dist
    from synthesize_pregrasp_b200 import _abi
    from synthesize_pregrasp_b200 import engine as EN
    from synthesize_pregrasp_b200.envspec import make_spec
    spec = make_spec(wl["env"])
    path, validate = _bench_path(spec)
    eng = EN.RolloutEngine(spec, device=local, validate=validate)
    strong = bool(wl.get("strong"))
    K = wl["K"] // world if strong else wl["K"]                # strong scaling: the named K is the whole job
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev).manual_seed(12367134 + rank)
    u = torch.zeros((N_STEPS, D), dtype=torch.float64, device=dev)
    noise = SIGMA * torch.randn((K, N_STEPS, D), generator=gen, device=dev, dtype=torch.float32)
    noise_host = noise.cpu().pin_memory()
    L = _abi.lib()
    Q = wl.get("Q", 0)
    if Q:
        uQ = torch.zeros((Q, N_STEPS, D), dtype=torch.float64, device=dev)
        probQ = torch.arange(Q, dtype=torch.int32, device=dev).repeat_interleave(K // Q).contiguous()
        pathsQ = np.tile(np.asarray(path, np.int32), (Q, 1))

    def step_device():
        if Q:
            J = eng.rollout_multi(uQ, noise, probQ, pathsQ, wl["force"])
            M = K // Q
            for q in range(Q):
                EN.mppi_update(J[q * M:(q + 1) * M], noise[q * M:(q + 1) * M], None, KAPPA, ALPHA, apply=False, want_partial=True)
            return J
        J = eng.rollout(u, noise, path, wl["force"])
        part = EN.mppi_update(J, noise, None, KAPPA, ALPHA, apply=False, want_partial=True)
        if world > 1:
            got = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(got, part)                          # the one exchange of an MPPI iteration: 99 doubles per rank
        return J

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # timing rule: inputs + per-rollout state larger than the 126 MB L2, or an L2 flush between the timed steps
    footprint_mb = K * (N_STEPS * D * 4 + 21408) / 1e6
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if footprint_mb <= 126 else None
    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = L.pg_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        if flush_buf is not None:
            flush_buf.fill_(i & 0xFF)            # 256 MB write = 2x L2; ~0.1 ms inside the timed region
        step_device()
        ev[i + 1].record()
    barrier()
    launches = int(L.pg_launch_count() - l0)
    total_ms = ev[0].elapsed_time(ev[-1])
    # dominant-kernel timing: physics and cost stages separately, CUDA events on the launching stream
    # (median of three samples: a single launch scatters by several per cent; one sample for the very large jobs)
    out = eng.physics(u, noise, path, wl["force"])
    samples = []
    for _ in range(3 if K <= (1 << 17) else 1):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record(); out = eng.physics(u, noise, path, wl["force"]); e1.record()
        eng.cost(out["pose"], out["fins"]); e2.record()
        torch.cuda.synchronize()
        samples.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    phys_ms = sorted(s[0] for s in samples)[len(samples) // 2]
    cost_ms = sorted(s[1] for s in samples)[len(samples) // 2]
    # e2e: host buffers in, host J out, through the plugin-facing C-ABI call
    nh = noise_host.numpy(); uh = np.zeros((N_STEPS, D))
    if K <= (1 << 17):
        eng.rollout_host(uh, nh, path, wl["force"])          # warm the staging buffers (very large jobs: the first call is the measurement)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(steps, 2)) if K <= (1 << 17) else 1
    for _ in range(e2e_steps):
        eng.rollout_host(uh, nh, path, wl["force"])
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sampler.stop_flag = True
    sampler.join(2)
    t = torch.tensor([total_ms, e2e_s * 1e3, phys_ms, cost_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, phys_ms, cost_ms = t.tolist()
    eng.close()
    if rank != 0:
        return None
    ms_per_step = total_ms / steps
    units = K * world * N_STEPS
    substeps = spec.substeps_per_rollout
    res = dict(workload=name, env=wl["env"], force=wl["force"], K=K, path=path, validate=validate, substeps=substeps, ms_per_step=ms_per_step,
               value=units / (ms_per_step * 1e-3), phys_ms=phys_ms, cost_ms=cost_ms, e2e_value=units / (e2e_ms * 1e-3),
               h2d=int(nh.nbytes + uh.nbytes), d2h=int(K * 8), launches=launches, clocks=sampler.summary(), footprint_mb=footprint_mb,
               flushed=flush_buf is not None, strong=strong)
    # bounded sample of the same workload on one host core (the oracle port; never the product)
    ref = CpuReference(wl["env"], wl["force"], 1, path=path, validate=validate)
    if cpu_sample:
        cpu_n = cpu_sample
        cpu_v, cpu_dt = ref.run(cpu_n)
    else:
        cpu_n, cpu_dt = 0, 0.0
        while cpu_dt < cpu_seconds:
            _, dt1 = ref.run(16, seed=cpu_n)
            cpu_n += 16; cpu_dt += dt1
        cpu_v = cpu_n * N_STEPS / cpu_dt
    res["cpu_baseline"] = {"value": cpu_v, "unit": "rollout-steps/s", "cores": 1, "kind": "port",
                           "sample": f"{cpu_n} rollouts of the same workload, 1 process, oracle/ (CPU restatement; Python-driven like the reference) in {cpu_dt:.1f} s"}
    return res


def rollout_line(args, r, world, peaks):
    """The contract's JSON object for one rollout workload."""
    FLOP_PER_EVAL = 164.9e6          # algorithmic FLOPs of one score evaluation at P=2048
    # audited FLOPs of the solver rows per substep ON THE BENCH PATH (tools/audit_substep_flops.py -> profiles/r2_flop_audit.json;
    # 1 FMA = 2, collision detection / integration not counted); SURVEY 8(d)'s 90 kFLOP estimate for workloads that were not audited
    audit_p = os.path.join(ROOT, "profiles", "r2_flop_audit.json")
    audit = json.load(open(audit_p)) if os.path.exists(audit_p) else {}
    a = audit.get(r["env"])
    FLOP_PER_SUBSTEP = a["solver_flop_per_substep"] if a and a.get("path") == list(r["path"]) else 90e3
    flop_src = "audited solver rows on this path, profiles/r2_flop_audit.json" if a and a.get("path") == list(r["path"]) else "SURVEY 8(d) estimate"
    K, phys_ms, cost_ms = r["K"], r["phys_ms"], r["cost_ms"]
    f64_peak, f64_src = _fp64_peak()
    bytes_phys = K * (N_STEPS * D * 4 + N_STEPS * (7 + 6) * 8 + 8) + N_STEPS * D * 8
    ach_hbm = bytes_phys / (phys_ms * 1e-3) / 1e9
    traffic = _ncu_traffic(r["workload"])
    roof = {"kernel": "rollout_kernel (physics, %.0f%% of the step)" % (100 * phys_ms / (phys_ms + cost_ms)), "bound": "hbm",
            "achieved": ach_hbm, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach_hbm / peaks["hbm"], "traffic": traffic,
            "peak_source": peaks["src"],
            "note": "algorithmic HBM bytes are ~400 B per rollout-step (u, noise in; pose, fins, metric out), so this fraction is tiny by "
                    "construction: the kernel is bound by instruction issue / dependent-instruction latency of its fp64 code "
                    "(roofline_fp64; ncu evidence in profiles/README.md).  traffic = dram bytes of one launch from an ncu capture of "
                    "this very build (profiles/ncu_traffic.json, matched by the sha256 of the library's sources or of the kernel's SASS), null when no such capture exists"}
    ach64 = FLOP_PER_SUBSTEP * r["substeps"] * K / (phys_ms * 1e-3) / 1e12
    roof64 = {"kernel": "rollout_kernel", "bound": "fp64-issue", "achieved": ach64, "peak": f64_peak, "unit": "TFLOP/s",
              "frac": ach64 / f64_peak, "peak_source": f64_src, "flop_per_substep": FLOP_PER_SUBSTEP,
              "flop_source": flop_src}
    ach_t = FLOP_PER_EVAL * K * N_STEPS / (cost_ms * 1e-3) / 1e12
    roof_cost = {"kernel": "pcn_tc_kernel (fused cost, tcgen05 3xTF32)", "bound": "tensor", "achieved": ach_t, "peak": peaks["bf16"],
                 "unit": "TFLOP/s", "frac": ach_t / peaks["bf16"], "peak_source": peaks["src"],
                 "note": "3x-split tf32 counts 1x algorithmic FLOP"}
    l2 = (f"inputs+state {r['footprint_mb']:.0f} MB > 126 MB L2" if not r["flushed"] else
          f"inputs+state {r['footprint_mb']:.0f} MB < 126 MB L2: flushed by a 256 MB write before every timed step (inside the timed region)")
    return {"metric": "rollout-steps/sec", "value": r["value"], "unit": "rollout-steps/s", "n_gpus": world, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if r["strong"] else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(r["workload"], r["env"], r["force"], K, r["path"], r["validate"], r["substeps"], l2),
            "substeps_per_s": K * world * r["substeps"] / (r["ms_per_step"] * 1e-3), "mppi_iters_per_s": 1e3 / r["ms_per_step"],
            "stage_ms": {"physics": phys_ms, "cost": cost_ms},
            "roofline": roof, "roofline_fp64": roof64, "roofline_cost": roof_cost, "cpu_baseline": r["cpu_baseline"],
            "e2e": {"value": r["e2e_value"], "unit": "rollout-steps/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"]},
            "gpu_launches": r["launches"], "clocks": r["clocks"], "host_cores": os.cpu_count() or 1, "lib_sha16": _lib_sha16(), "src_sha16": _src_sha16()}


def _config(workload, env, force, K, path, validate, substeps, l2=None):
    c = {"workload": workload, "env": env, "max_force": force, "K_per_gpu": K, "N": N_STEPS, "D": D, "substeps_per_rollout": substeps,
         "sigma": SIGMA, "path": path, "contact_graph": "csg.npy / paths_id.npy[0]" if validate else "dummy_states.npy"}
    if l2 is not None:
        c["l2"] = l2
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bookshelf_K65536")
    ap.add_argument("--K", type=int, default=None, help="override rollouts per GPU (profiling runs)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="rollouts in the CPU baseline sample")
    ap.add_argument("--no-extra", action="store_true", help="skip the plate lines attached to the default workload")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    cores = os.cpu_count() or 1
    if args.workload in OTHER_WORKLOADS:
        return OTHER_WORKLOADS[args.workload](args, rank, world, local)
    wl = dict(WORKLOADS[args.workload])
    if args.K:
        wl["K"] = args.K

    if args.impl == "reference":
        if rank != 0:
            return
        from synthesize_pregrasp_b200.envspec import make_spec
        spec = make_spec(wl["env"])
        path, validate = _bench_path(spec)
        threads = max(1, cores)
        # a bounded sample of the arm's workload per step: a power of two, >= 4096 on boxes with enough cores to finish a step in ~10 s
        n = args.cpu_sample or max(256, min(4096, 256 * threads))
        ref = CpuReference(wl["env"], wl["force"], threads, path=path, validate=validate)
        vals = []
        for i in range(args.warmup + args.steps):
            if i == 0 and args.warmup > 0:
                v, dt = ref.run(max(threads, n // 8), seed=i)     # warm-up steps are shorter: the pool is already warm
            else:
                v, dt = ref.run(n if i >= args.warmup else max(threads, n // 8), seed=i)
            if i >= args.warmup:
                vals.append((v, dt))
        ref.close()
        v = float(np.mean([x[0] for x in vals])); ms = float(np.mean([x[1] for x in vals])) * 1e3
        sample = (f"{n} rollouts of the {args.workload} workload per step on {threads} host processes (oracle/, CPU restatement of the "
                  "PyBullet path; PyBullet itself is not installable here); the rate is per rollout-step, independent of the sample size")
        print(json.dumps({"metric": "rollout-steps/sec", "value": v, "unit": "rollout-steps/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "impl": "reference",
                          "config": _config(args.workload, wl["env"], wl["force"], wl["K"], path, validate, spec.substeps_per_rollout),
                          "cpu_baseline": {"value": v, "unit": "rollout-steps/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": v, "unit": "rollout-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = _peaks()
    r = measure_rollouts(args, args.workload, wl, args.steps, args.warmup, rank, world, local, cpu_seconds=12.0, cpu_sample=args.cpu_sample)
    extras = []
    if args.workload == "bookshelf_K65536" and not args.K and not args.no_extra:
        # the task the north-star target is stated on, inside the driver's own run: plate at the headline batch size and at the
        # reference's own 12 x 30 samples, each with its CPU-port baseline
        for name, st, wu in (("plate_K65536", 3, 2), ("plate_K360", 5, 3)):
            x = measure_rollouts(args, name, dict(WORKLOADS[name]), st, wu, rank, world, local, cpu_seconds=6.0)
            if x is not None:
                x["steps"], x["warmup"] = st, wu
                extras.append(rollout_line(args, x, world, peaks))
    if rank == 0:
        r["steps"], r["warmup"] = args.steps, args.warmup
        line = rollout_line(args, r, world, peaks)
        if extras:
            line["extra"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
