"""Thin host layer over the C ABI: owns the env handle, moves torch tensors' device pointers across.

torch is plumbing here (device memory, streams, torch.distributed); all arithmetic of the hot path runs in
the hand-written kernels of libpregrasp_b200.so.  There is no CPU fallback: without a CUDA device or without
the library every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from . import envspec as ES

ACTION_DIM = ES.ACTION_DIM


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pregrasp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class RolloutEngine:
    """One environment on one GPU: replaces `env_fn(render=False, device=..., **kwargs)` plus the serial
    rollout loop of the reference's worker process (stoch_traj_opt.py:113-136)."""

    def __init__(self, env: str | ES.EnvSpec, states=None, device: int | None = None, validate: bool = False):
        _require_cuda()
        self.lib = _abi.lib()
        self.spec = env if isinstance(env, ES.EnvSpec) else ES.make_spec(env)
        if states is None:
            states = self.spec.csg_states if validate else self.spec.dummy_states
        self.states = np.asarray(states)
        self.device = torch.cuda.current_device() if device is None else int(device)
        cspec, self._keep = _abi.make_c_spec(self.spec, self.states)
        h = C.c_void_p()
        _abi.check(self.lib.pg_env_create(C.byref(cspec), self.device, C.byref(h)))
        self.handle = h

    def set_rollout_mode(self, mode: int):
        """0 auto, 1 persistent one-launch kernel, 2 phased with work-sorted rollouts (pg_set_rollout_mode); same results."""
        _abi.check(self.lib.pg_set_rollout_mode(self.handle, int(mode)))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pg_env_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _dev(self):
        return torch.device("cuda", self.device)

    def _path(self, path, N):
        p = np.ascontiguousarray(np.asarray(path, np.int32)[:N])
        if len(p) < N:
            raise ValueError(f"path has {len(p)} states, need {N}")
        return p

    def rollout(self, u, noise, path, max_force, K=None, want_metric=False, want_trace=False):
        """u [N,48] (any array, copied to the device as f64); noise [K,N,48] float32 CUDA tensor or None
        (then K copies of u).  -> J [K] f64 CUDA tensor (+ metric [K], + trace [K,N*50,7])."""
        with torch.cuda.device(self.device):
            u_t = torch.as_tensor(np.asarray(u, np.float64) if not torch.is_tensor(u) else u, dtype=torch.float64,
                                  device=self._dev()).contiguous()
            N = u_t.shape[0]
            if noise is not None:
                assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous()
                K = noise.shape[0]
                assert tuple(noise.shape[1:]) == (N, ACTION_DIM)
            elif K is None:
                K = 1
            J = torch.empty(K, dtype=torch.float64, device=self._dev())
            metric = torch.empty(K, dtype=torch.float64, device=self._dev()) if want_metric else None
            trace = torch.empty((K, N * ES.CONTROL_SKIP, 7), dtype=torch.float64, device=self._dev()) if want_trace else None
            p = self._path(path, N)
            _abi.check(self.lib.pg_rollout(self.handle, _ptr(u_t), _ptr(noise), K, N, p.ctypes.data_as(_abi._IP),
                                           float(max_force), _ptr(J), _ptr(metric), _ptr(trace), _stream()))
        out = [J]
        if want_metric:
            out.append(metric)
        if want_trace:
            out.append(trace)
        return out[0] if len(out) == 1 else tuple(out)

    def rollout_multi(self, u, noise, prob, paths, max_force, want_metric=False):
        """Q independent MPPI problems in one launch (model_optimize.py:74-117 runs them one after the other).
        u [Q,N,48] f64 CUDA tensor, noise [K,N,48] f32 CUDA tensor, prob [K] int32 CUDA tensor (problem of each rollout),
        paths [Q,N] contact states.  -> J [K] (+ metric [K])."""
        with torch.cuda.device(self.device):
            assert u.is_cuda and u.dtype == torch.float64 and u.is_contiguous() and u.dim() == 3
            assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous()
            assert prob.is_cuda and prob.dtype == torch.int32 and prob.is_contiguous()
            Q, N = u.shape[0], u.shape[1]
            K = noise.shape[0]
            assert tuple(noise.shape[1:]) == (N, ACTION_DIM) and prob.shape[0] == K
            p = np.ascontiguousarray(np.asarray(paths, np.int32)[:, :N])
            assert p.shape == (Q, N)
            J = torch.empty(K, dtype=torch.float64, device=self._dev())
            metric = torch.empty(K, dtype=torch.float64, device=self._dev()) if want_metric else None
            _abi.check(self.lib.pg_rollout_multi(self.handle, _ptr(u), _ptr(noise), K, N, Q, _ptr(prob), p.ctypes.data_as(_abi._IP),
                                                 float(max_force), _ptr(J), _ptr(metric), _stream()))
        return (J, metric) if want_metric else J

    def rollout_host(self, u, noise, path, max_force, K=None, want_metric=False, want_trace=False):
        """Host-buffer form (numpy in, numpy out; H2D/D2H inside) -- the plugin-facing call."""
        u = np.ascontiguousarray(u, np.float64)
        N = u.shape[0]
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float32)
            K = noise.shape[0]
        elif K is None:
            K = 1
        J = np.empty(K, np.float64)
        metric = np.empty(K, np.float64) if want_metric else None
        trace = np.empty((K, N * ES.CONTROL_SKIP, 7), np.float64) if want_trace else None
        p = self._path(path, N)
        dp = lambda a: a.ctypes.data_as(_abi._DP) if a is not None else None
        _abi.check(self.lib.pg_rollout_host(self.handle, dp(u), noise.ctypes.data_as(_abi._FP) if noise is not None else None,
                                            K, N, p.ctypes.data_as(_abi._IP), float(max_force), dp(J), dp(metric), dp(trace)))
        out = [J]
        if want_metric:
            out.append(metric)
        if want_trace:
            out.append(trace)
        return out[0] if len(out) == 1 else tuple(out)

    def physics(self, u, noise, path, max_force, K=None, want_trace=False, want_iters=False):
        with torch.cuda.device(self.device):
            u_t = torch.as_tensor(np.asarray(u, np.float64) if not torch.is_tensor(u) else u, dtype=torch.float64,
                                  device=self._dev()).contiguous()
            N = u_t.shape[0]
            if noise is not None:
                K = noise.shape[0]
            elif K is None:
                K = 1
            d = self._dev()
            pose = torch.empty((K, N, 7), dtype=torch.float64, device=d)
            fins = torch.empty((K, N, 2, 3), dtype=torch.float64, device=d)
            metric = torch.empty(K, dtype=torch.float64, device=d)
            trace = torch.empty((K, N * ES.CONTROL_SKIP, 7), dtype=torch.float64, device=d) if want_trace else None
            iters = torch.empty((K, N * ES.CONTROL_SKIP), dtype=torch.int32, device=d) if want_iters else None
            p = self._path(path, N)
            _abi.check(self.lib.pg_physics(self.handle, _ptr(u_t), _ptr(noise), K, N, p.ctypes.data_as(_abi._IP),
                                           float(max_force), _ptr(pose), _ptr(fins), _ptr(metric), _ptr(trace), _ptr(iters),
                                           _stream()))
        return dict(pose=pose, fins=fins, metric=metric, trace=trace, iters=iters)

    def cost(self, pose, fins):
        """pose [B,7] f64, fins [B,2,3] f64 CUDA tensors -> score logits [B] f32 (reward = 100 x logit)."""
        with torch.cuda.device(self.device):
            pose = pose.reshape(-1, 7).contiguous()
            fins = fins.reshape(-1, 6).contiguous()
            B = pose.shape[0]
            logit = torch.empty(B, dtype=torch.float32, device=self._dev())
            _abi.check(self.lib.pg_cost(self.handle, _ptr(pose), _ptr(fins), B, _ptr(logit), _stream()))
        return logit

    def score(self, pts, df, grasp, mask=None):
        """LargeScoreFunction.pred_score on explicit clouds: pts [B,P,3] f32, df [B,P] f32, grasp [B,6] f32."""
        with torch.cuda.device(self.device):
            pts, df, grasp = pts.contiguous(), df.contiguous(), grasp.contiguous()
            B, P = pts.shape[0], pts.shape[1]
            if mask is not None:
                mask = mask.to(torch.uint8).contiguous()
            logit = torch.empty(B, dtype=torch.float32, device=self._dev())
            _abi.check(self.lib.pg_score(self.handle, _ptr(pts), _ptr(df), _ptr(mask), _ptr(grasp), B, P, _ptr(logit), _stream()))
        return logit


class Episode:
    """One episode advanced one env step per call (pg_env_step_single): the gym-style reset()/step() of the reference envs
    (plate_contact_graph_env.py:176-250, :253-632).  The state between steps (PgCarry) stays on the GPU."""

    def __init__(self, engine: RolloutEngine):
        self.engine, self.lib = engine, engine.lib
        h = C.c_void_p()
        _abi.check(self.lib.pg_carry_create(engine.handle, C.byref(h)))
        self.handle = h

    def reset(self):
        pose = np.zeros(7)
        _abi.check(self.lib.pg_carry_reset(self.handle, pose.ctypes.data_as(_abi._DP)))
        return pose

    def step(self, a, state, max_force, last_step, want_trace=False):
        """-> (reward, metric, pose[7], trace[50,7] or None)"""
        a = np.ascontiguousarray(a, np.float64)
        assert a.shape == (ACTION_DIM,)
        reward, metric, pose = np.zeros(1), np.zeros(1), np.zeros(7)
        trace = np.zeros((ES.CONTROL_SKIP, 7)) if want_trace else None
        dp = lambda x: x.ctypes.data_as(_abi._DP) if x is not None else None
        _abi.check(self.lib.pg_env_step_single(self.engine.handle, self.handle, dp(a), int(state), float(max_force), 1 if last_step else 0,
                                               dp(reward), dp(metric), dp(pose), dp(trace)))
        return float(reward[0]), float(metric[0]), pose, trace

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pg_carry_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mppi_update(J, noise, u, kappa=5.0, alpha=1.0, apply=True, want_partial=False):
    """Exp-weighted update on the device (stoch_traj_opt.py:195-204).  J [K] f64, noise [K,N,D] f32, u [N,D] f64
    CUDA tensors; u is updated in place when apply.  -> partial (min J, sum w, sum J, sum w*noise) if asked."""
    _require_cuda()
    L = _abi.lib()
    K, N, D = noise.shape
    partial = torch.empty(3 + N * D, dtype=torch.float64, device=J.device) if want_partial else None
    with torch.cuda.device(J.device):
        _abi.check(L.pg_mppi_update(_ptr(J), _ptr(noise), K, N, D, float(kappa), float(alpha), _ptr(u), _ptr(partial),
                                    1 if apply else 0, _stream()))
    return partial


def mppi_merge_apply(partials, u, kappa=5.0, alpha=1.0):
    """Host merge of G per-rank partials [G,3+N*D] (log-sum-exp rescale to the global min) and update of u [N,D]."""
    L = _abi.lib()
    P = np.ascontiguousarray(partials, np.float64)
    u = np.ascontiguousarray(u, np.float64)
    N, D = u.shape
    _abi.check(L.pg_mppi_merge_apply(P.ctypes.data_as(_abi._DP), P.shape[0], N, D, float(kappa), float(alpha),
                                     u.ctypes.data_as(_abi._DP)))
    return u
