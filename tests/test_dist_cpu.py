"""world_size-2 gloo test of the per-iteration exchange: each rank reduces its shard to one partial vector,
one all_gather, identical merged update on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import mppi
    from synthesize_pregrasp_b200.engine import mppi_merge_apply
    rng = np.random.RandomState(11)
    K, N, D = 101, 2, 48
    J = rng.normal(700, 30, K); E = (0.8 * rng.randn(K, N, D)).astype(np.float32); u = rng.randn(N, D)
    lo, hi = (0, 57) if rank == 0 else (57, K)
    m, s, v, sj = mppi.partials(J[lo:hi], E[lo:hi], 5.0)          # what pg_mppi_update writes per rank
    part = torch.tensor(np.r_[m, s, sj, v.ravel()])
    got = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(got, part)
    unew = mppi_merge_apply(torch.stack(got).numpy(), u.copy(), 5.0, 1.0)
    ref, _ = mppi.mppi_update(u, J, E, 5.0, 1.0)
    q.put((rank, float(np.abs(unew - ref).max())))
    dist.destroy_process_group()


def test_two_rank_merge_matches_single_process_update():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert all(e <= 1e-13 for _, e in res), res


def test_triple_index_space_partition():
    """The feasibility sweeps shard the C(P,3) triple index space across ranks: contiguous, disjoint, covering."""
    from synthesize_pregrasp_b200.dyn_feasibility_check import triple_range
    for P in (3, 14, 256, 1024):
        total = P * (P - 1) * (P - 2) // 6
        for world in (1, 2, 3, 8):
            edges = [triple_range(r, world, P) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (a0, a1), (b0, b1) in zip(edges[:-1], edges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
