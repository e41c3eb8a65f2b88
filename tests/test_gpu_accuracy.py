"""Who is right when oracle and CUDA path disagree?  On the ill-conditioned c2 problem
(cond(K) ~ 5e7 for objective 1) they differ by ~1e-8 * |intercepts| because LAPACK and the CUDA
Cholesky / solves round differently.  Here the GP part is recomputed in 80-bit extended precision
(numpy longdouble, eps 1e-19) as the truth: the CUDA path must be at least about as close to it
as the float64 oracle is."""
import math

import numpy as np
import pytest
import torch

from helpers import oracle_model
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu
LD = np.longdouble


def _matern_ld(x1, x2, ls, s):
    a = x1.astype(LD) / ls.astype(LD)
    b = x2.astype(LD) / ls.astype(LD)
    sq = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    r = np.sqrt(np.maximum(sq, LD(1e-30)))
    s5 = np.sqrt(LD(5.0))
    return LD(s) * (LD(1) + s5 * r + LD(5) / LD(3) * r * r) * np.exp(-s5 * r)


def _chol_ld(K):
    n = K.shape[0]
    L = np.zeros_like(K)
    for j in range(n):
        L[j, j] = np.sqrt(K[j, j] - (L[j, :j] ** 2).sum())
        if j + 1 < n:
            L[j + 1:, j] = (K[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
    return L


def _solve_ld(L, R):
    n = L.shape[0]
    Y = np.zeros_like(R)
    for i in range(n):
        Y[i] = (R[i] - L[i, :i] @ Y[:i]) / L[i, i]
    X = np.zeros_like(R)
    for i in range(n - 1, -1, -1):
        X[i] = (Y[i] - L[i + 1:, i] @ X[i + 1:]) / L[i, i]
    return X


def test_cuda_path_is_as_accurate_as_the_float64_oracle():
    if np.finfo(LD).eps > 1e-18:
        pytest.skip("no extended-precision long double on this platform")
    from decoupledbo_b200 import synthetic
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    P = synthetic.problem_c2(n_cand=6)
    om = oracle_model(P.model, distance="direct")
    xd = P.x_disc.numpy()
    W = P.weights
    # extended-precision GP state per objective
    st = []
    for o in P.model.models:
        x = o.train_x.numpy()
        ls = o.lengthscale.numpy()
        K = _matern_ld(x, x, ls, o.outputscale) + LD(o.noise) * np.eye(o.n, dtype=LD)
        L = _chol_ld(K)
        alpha = _solve_ld(L, (o.train_y.numpy().astype(LD) - LD(o.mean_const))[:, None])[:, 0]
        st.append((x, ls, L, alpha, o))
    for target in (0, 1):
        acq = DiscreteKnowledgeGradient(P.model, P.x_disc, W, target_output_ix=target)
        with torch.no_grad():
            kg_gpu = acq(P.candidates.unsqueeze(1)).numpy()
        err_gpu, err_or = [], []
        for c in range(P.candidates.shape[0]):
            xc = P.candidates[c].numpy()
            pts = np.concatenate([xc[None], xd])
            means = []
            for (x, ls, L, alpha, o) in st:
                means.append(LD(o.mean_const) + _matern_ld(pts, x, ls, o.outputscale) @ alpha)
            x, ls, L, alpha, o = st[target]
            kx = _matern_ld(xc[None], x, ls, o.outputscale)[0]
            sol = _solve_ld(L, kx[:, None])[:, 0]
            cov = _matern_ld(xc[None], pts, ls, o.outputscale)[0] - _matern_ld(pts, x, ls, o.outputscale) @ sol
            z = cov / np.sqrt(cov[0] + LD(o.noise))
            a = (np.stack(means, -1) @ W.numpy().astype(LD).T).T  # (S, N+1)
            b = W[:, target].numpy().astype(LD)[:, None] * z[None, :]
            truth = np.mean([
                odk.expected_max_gradients_np(a[j].astype(np.float64), b[j].astype(np.float64))[0]
                - float(a[j].astype(np.float64).max()) for j in range(W.shape[0])])
            kg_or = odk.kg_single_output(om, P.candidates[c], target, P.x_disc, W, dense=False).item()
            err_gpu.append(abs(kg_gpu[c] - truth))
            err_or.append(abs(kg_or - truth))
        scale = float(torch.cat([o.train_y for o in P.model.models]).abs().max())
        # both float64 paths sit at the conditioning floor; the CUDA path must not be worse by more
        # than a small factor (and both must be far below the values themselves)
        assert max(err_gpu) <= 8.0 * max(max(err_or), 1e-13 * scale), (err_gpu, err_or)
        assert max(err_gpu) <= 1e-6 * scale
