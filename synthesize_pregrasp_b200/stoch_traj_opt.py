"""StochTrajOptimizer with the reference's constructor / optimize() / replay_traj() contract
(stoch_traj_opt.py:31-46, :155-231, :233-247), running every rollout of an MPPI iteration in one batched
GPU call instead of Num_processes PyBullet worker processes.

Differences that are deliberate and documented: noise comes from a seeded torch generator (the
reference reseeds numpy with the worker pid, so its noise is not reproducible); with torch.distributed
initialised each rank simulates K/world_size samples and the ranks exchange ONE small vector per iteration
(min J, sum w, sum J, sum w*E) instead of pickled (J, E) arrays over pipes.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import engine as EN


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


class StochTrajOptimizer:
    def __init__(self, env, sigma=0.2, kappa=5, alpha=1, Num_processes=64, Traj_per_process=15, TimeSteps=200,
                 Iterations=20, seed=None, render=True, initial_guess=None, verbose=0, patience=10, num_gpus=4, **kwargs):
        self.sigma, self.kappa, self.alpha = sigma, kappa, alpha
        self.num_process, self.Traj_per_process = Num_processes, Traj_per_process
        self.N, self.ITER = TimeSteps, Iterations
        self.M = Num_processes * Traj_per_process              # total number of trajectories simulated
        self.seed, self.render, self.initial_guess = seed, render, initial_guess
        self.env, self.verbose, self.patience, self.kwargs = env, verbose, patience, kwargs
        self.world = None

    def reset(self):
        # the "replay" env of the reference (stoch_traj_opt.py:86-88) doubles as the batched engine here
        self.world = self.env(render=False, device="cuda", **self.kwargs)
        self.ctrl_dim = self.world.action_space.shape[0]
        d = _dist()
        self.rank, self.world_size = (d.get_rank(), d.get_world_size()) if d else (0, 1)
        self.K_local = self.M // self.world_size + (1 if self.rank < self.M % self.world_size else 0)
        self.gen = torch.Generator(device=self.world.engine._dev())
        self.gen.manual_seed((self.seed if self.seed is not None else 0) + self.rank)

    def _rollouts(self, u):
        eng = self.world.engine
        noise = torch.randn((self.K_local, self.N, self.ctrl_dim), generator=self.gen, device=eng._dev(),
                            dtype=torch.float32) * float(self.sigma)
        J = eng.rollout(u, noise, self.world.path, self.world.max_forces)
        return J, noise

    def optimize(self):
        self.reset()
        u = np.load(self.initial_guess) if self.initial_guess is not None else np.zeros((self.N, self.ctrl_dim))
        u = np.array(u, np.float64)
        self.uopt, self.Jopt, self.metric_opt = u.copy(), np.inf, -np.inf
        Jopt_log, r_log = [], []
        dev = self.world.engine._dev()
        d = _dist()
        start_time = time.time()
        for iter_id in range(self.ITER):
            J, noise = self._rollouts(u)
            if d is None:
                u_t = torch.as_tensor(u, dtype=torch.float64, device=dev).contiguous()
                part = EN.mppi_update(J, noise, u_t, self.kappa, self.alpha, apply=True, want_partial=True)
                u = u_t.cpu().numpy()
                Jmean = float(part[2].item()) / self.M
            else:
                part = EN.mppi_update(J, noise, None, self.kappa, self.alpha, apply=False, want_partial=True)
                gathered = [torch.empty_like(part) for _ in range(self.world_size)]
                d.all_gather(gathered, part)                      # the single exchange of the iteration
                P = torch.stack(gathered).cpu().numpy()
                u = EN.mppi_merge_apply(P, u, self.kappa, self.alpha)
                Jmean = float(P[:, 2].sum()) / self.M
            end_time = time.time()
            Jcur, metric = self.replay_traj(u)
            if Jcur <= self.Jopt:
                self.uopt, self.Jopt = u.copy(), Jcur
            if metric >= self.metric_opt:
                self.metric_opt = metric
            if self.verbose == 1 and self.rank == 0:
                print(f"Iter {iter_id + 1} took {round(end_time - start_time, 2)} sec (mean R: {round(-Jmean, 2)}).",
                      f"Cur R: {round(float(-Jcur), 2)}, Opt R {round(float(-self.Jopt), 2)}",
                      f"Cur M: {round(metric, 2)} Opt M: {round(self.metric_opt, 2)}")
            Jopt_log.append(-self.Jopt)
            r_log.append(-Jcur)
            start_time = time.time()
        return self.uopt, -self.Jopt, Jopt_log, r_log

    def replay_traj(self, u):
        J, metric = self.world.engine.rollout(u, None, self.world.path, self.world.max_forces, K=1, want_metric=True)
        return float(J.item()), float(metric.item())

    def render_traj(self, u, replay_times=10):
        raise NotImplementedError("rendering is out of scope of the GPU rollout path (SURVEY.md section 2, row 17)")


class MultiStochTrajOptimizer:
    """Q independent StochTrajOptimizer problems advanced in lock-step, one launch per MPPI iteration.

    model_optimize.py:74-117 builds one optimiser per contact-state path (up to 10) and per repetition (`--runs`, 3) and runs
    them one after the other, each with only Num_processes*Traj_per_process = 360 samples -- far too few to fill a GPU with
    one thread per rollout.  The problems share the env and differ only in path, seed and control iterate, so here every
    iteration is ONE `pg_rollout_multi` launch over all Q*M rollouts, Q segment-wise exp-weighted updates and one replay
    launch of the Q updated iterates.  Problem q reproduces a single StochTrajOptimizer with seed `seeds[q]` exactly
    (same generator stream, same update kernel).
    """

    def __init__(self, env, paths, seeds, sigma=0.2, kappa=5, alpha=1, Num_processes=64, Traj_per_process=15, TimeSteps=200,
                 Iterations=20, verbose=0, **kwargs):
        assert len(paths) == len(seeds)
        self.env, self.paths, self.seeds = env, [list(p) for p in paths], list(seeds)
        self.sigma, self.kappa, self.alpha = sigma, kappa, alpha
        self.M = Num_processes * Traj_per_process
        self.N, self.ITER, self.verbose, self.kwargs = TimeSteps, Iterations, verbose, kwargs

    def optimize(self):
        Q, M, N = len(self.paths), self.M, self.N
        world = self.env(render=False, device="cuda", **self.kwargs)
        eng, D = world.engine, world.action_space.shape[0]
        dev = eng._dev()
        gens = []
        for s in self.seeds:
            g = torch.Generator(device=dev)
            g.manual_seed(s)
            gens.append(g)
        paths = np.asarray([p[:N] for p in self.paths], np.int32)
        prob = torch.arange(Q, dtype=torch.int32, device=dev).repeat_interleave(M).contiguous()
        prob1 = torch.arange(Q, dtype=torch.int32, device=dev).contiguous()
        u = torch.zeros((Q, N, D), dtype=torch.float64, device=dev)
        uopt = u.clone()
        Jopt = np.full(Q, np.inf)
        logs = [[] for _ in range(Q)]
        zero_noise = torch.zeros((Q, N, D), dtype=torch.float32, device=dev)
        for it in range(self.ITER):
            noise = torch.cat([torch.randn((M, N, D), generator=g, device=dev, dtype=torch.float32) * float(self.sigma) for g in gens])
            J = eng.rollout_multi(u, noise, prob, paths, world.max_forces)
            for q in range(Q):
                EN.mppi_update(J[q * M:(q + 1) * M], noise[q * M:(q + 1) * M], u[q], self.kappa, self.alpha, apply=True)
            Jcur = eng.rollout_multi(u, zero_noise, prob1, paths, world.max_forces).cpu().numpy()      # replay_traj of every problem
            for q in range(Q):
                if Jcur[q] <= Jopt[q]:
                    Jopt[q] = Jcur[q]
                    uopt[q] = u[q]
                logs[q].append(-Jopt[q])
            if self.verbose:
                print(f"Iter {it + 1}: best R per problem {np.round(-Jopt, 2).tolist()}")
        return uopt.cpu().numpy(), -Jopt, logs
