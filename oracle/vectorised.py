"""ORACLE / CPU baseline "(V)" of BASELINE.md section 3 -- TEST AND BENCH INFRASTRUCTURE ONLY.

A vectorised, row-only, batched CPU implementation of the decoupled discrete KG (the same
mathematics as ``discretekg.py:238-338``), written to be the STRONG host baseline next to the
reference-faithful one (per-candidate loop, dense (N+1)^2 covariance, Python march over all lines):

* everything that does not depend on the candidate is computed once (``prepare``): posterior means at
  the discretisation, ``K^-1 k(X_train, X_disc)``;
* a whole batch of candidates goes through one GEMM for the covariance rows (torch float64, all
  host threads);
* per (candidate, scalarisation) the lines are first cut down by the 3-point chord test (a line
  strictly below the chain min-slope -> max-intercept -> max-slope cannot be on the upper envelope)
  in vectorised numpy, then the reference's exact march (``oracle.discretekg.epigraph_march_np``)
  runs on the few survivors;
* the gradient uses the envelope theorem (``dE/da_k = dPhi_k``, ``dE/db_k = -dphi_k`` on hull lines)
  and ONE autograd pass through the batched GP part.

Only ``tests/`` and ``bench.py``'s cpu_baseline / reference legs may import this module; the product
path never does.  Results equal ``oracle.discretekg.forward`` to rounding (tests/test_oracle_vectorised.py).
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np
import torch
from torch import Tensor

from . import discretekg as odk
from . import gp as ogp

_EPS128 = 128 * 2.0**-52


class Prepared:
    """Candidate-independent state for one target objective."""

    def __init__(self, model: ogp.OracleModelList, discretisation: Tensor, weights: Tensor, target: int):
        self.model, self.disc, self.W, self.target = model, discretisation.double(), weights.double(), int(target)
        with torch.no_grad():
            mu = []
            for o in model.models:
                c = ogp._train_cache(o)
                mu.append((o.mean_const + ogp.kernel_matrix(o, self.disc, o.train_x) @ c["mean_cache"]) * o.y_std + o.y_mean)
            self.mu_disc = torch.stack(mu, dim=-1)  # (N, M)
            self.A0 = (self.W @ self.mu_disc.T).contiguous()  # (S, N)
            ot = model.models[self.target]
            ct = ogp._train_cache(ot)
            self.Bmat = torch.cholesky_solve(ogp.kernel_matrix(ot, ot.train_x, self.disc), ct["L"])  # (n, N)
        self.A0_np = self.A0.numpy()
        self.A0_max = self.A0_np.max(axis=1)


def _lines_batch(P: Prepared, X: Tensor):
    """Differentiable slope rows z (B, N+1) and own-line intercepts (B, S); own line FIRST, as the
    reference orders them (discretekg.py:277)."""
    model, ot = P.model, P.model.models[P.target]
    ct = ogp._train_cache(ot)
    mu_x = []
    for o in model.models:
        c = ogp._train_cache(o)
        mu_x.append((o.mean_const + ogp.kernel_matrix(o, X, o.train_x) @ c["mean_cache"]) * o.y_std + o.y_mean)
    mu_x = torch.stack(mu_x, dim=-1)  # (B, M)
    a_own = mu_x @ P.W.T  # (B, S)
    kx = ogp.kernel_matrix(ot, X, ot.train_x)  # (B, n)
    sol = torch.cholesky_solve(kx.T, ct["L"])  # (n, B)
    s2 = ot.y_std**2
    cov_disc = (ogp.kernel_matrix(ot, X, P.disc) - kx @ P.Bmat) * s2  # (B, N)
    cov_own = (ot.outputscale - (kx * sol.T).sum(-1)) * s2  # k(x, x) = outputscale for stationary kernels
    var = cov_own + ot.noise * s2
    z = torch.cat([cov_own.unsqueeze(1), cov_disc], dim=1) / var.sqrt().unsqueeze(1)
    return z, a_own


def _set_hull(a: np.ndarray, b: np.ndarray):
    """Exact hull of one line set through the chord prefilter: (E, idx, dE/da[h], dE/db[h])."""
    if np.all(np.abs(b) < odk.SLOPE_SHORTCUT_TOL):
        E, idx, dp, dq, _ = odk.expected_max_gradients_np(a, b)
        return E, idx, dp, dq
    iP, iQ, iT = int(np.argmin(b)), int(np.argmax(b)), int(np.argmax(a))
    bP, bQ, bT, aP, aQ, aT = b[iP], b[iQ], b[iT], a[iP], a[iQ], a[iT]
    keep = np.zeros(a.shape[0], dtype=bool)
    t = np.full(a.shape[0], np.inf)
    if bT > bP:
        m1 = (aT - aP) / (bT - bP)
        slack = _EPS128 * (abs(aT) + abs(aP) + abs(m1) * max(abs(bP), abs(bT)))
        t = np.minimum(t, aP + m1 * (b - bP) - slack)
    if bQ > bT:
        m2 = (aQ - aT) / (bQ - bT)
        slack = _EPS128 * (abs(aT) + abs(aQ) + abs(m2) * max(abs(bQ), abs(bT)))
        t = np.minimum(t, aT + m2 * (b - bT) - slack)
    keep = a > t
    keep[[iP, iQ, iT, 0]] = True
    sub = np.nonzero(keep)[0]
    E, idx, dp, dq, _ = odk.expected_max_gradients_np(a[sub], b[sub])
    return E, sub[idx], dp, dq


def kg_batch(P: Prepared, X: Tensor, need_grad: bool = True, threads: Optional[int] = None):
    """KG values (B,) and, if ``need_grad``, dKG/dX (B, d) for a batch of candidates (B, d)."""
    X = X.detach().double().clone().requires_grad_(need_grad)
    with torch.set_grad_enabled(need_grad):
        z, a_own = _lines_batch(P, X)
    B, S = a_own.shape
    z_np, ao_np = z.detach().numpy(), a_own.detach().numpy()
    w_t = P.W[:, P.target].numpy()
    Gz = np.zeros_like(z_np)
    Ga = np.zeros_like(ao_np)
    kg = np.zeros(B)

    def one_row(bi):
        a = np.empty(z_np.shape[1])
        acc = 0.0
        for j in range(S):
            a[0] = ao_np[bi, j]
            a[1:] = P.A0_np[j]
            bj = w_t[j] * z_np[bi]
            E, idx, dp, dq = _set_hull(a, bj)
            amax_own = a[0] >= P.A0_max[j]
            acc += E - (a[0] if amax_own else P.A0_max[j])
            if need_grad:
                np.add.at(Gz[bi], idx, dq * (w_t[j] / S))
                own = idx == 0
                Ga[bi, j] = (dp[own].sum() - (1.0 if amax_own else 0.0)) / S
        kg[bi] = acc / S

    if threads is None or threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(one_row, range(B)))
    else:
        for bi in range(B):
            one_row(bi)
    dX = None
    if need_grad:
        surrogate = (torch.from_numpy(Gz) * z).sum() + (torch.from_numpy(Ga) * a_own).sum()
        (dX,) = torch.autograd.grad(surrogate, X)
    return torch.from_numpy(kg), dX
