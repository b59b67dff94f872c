"""MPPI exp-weighted update (test oracle): stoch_traj_opt.py:192-204 verbatim semantics.
E is [K,N,D] here (the reference holds [N,D,K]); u is updated as u[j,k] += alpha*sum_i S_i E[i,j,k]."""
import numpy as np


def mppi_update(u, J, E, kappa=5.0, alpha=1.0):
    J = np.asarray(J, np.float64)
    J = J - J.min()                         # :195
    S = np.exp(-kappa * J)                  # :196
    S = S / S.sum()                         # :197-198
    un = np.array(u, np.float64, copy=True)
    N, D = un.shape
    for j in range(N):                      # :202-204
        for k in range(D):
            un[j, k] = un[j, k] + alpha * np.sum(S * np.asarray(E, np.float64)[:, j, k])
    return un, S


def partials(J, E, kappa=5.0):
    """Per-shard partial (m, s, v, sumJ) of SURVEY 8(e): m=min J, s=sum exp(-k(J-m)), v=sum exp(-k(J-m)) E."""
    J = np.asarray(J, np.float64)
    m = J.min()
    wgt = np.exp(-kappa * (J - m))
    return m, wgt.sum(), np.tensordot(wgt, np.asarray(E, np.float64), axes=(0, 0)), J.sum()


def merge_partials(parts, kappa=5.0):
    m = min(p[0] for p in parts)
    s = sum(p[1] * np.exp(-kappa * (p[0] - m)) for p in parts)
    v = sum(p[2] * np.exp(-kappa * (p[0] - m)) for p in parts)
    return m, s, v, sum(p[3] for p in parts)
