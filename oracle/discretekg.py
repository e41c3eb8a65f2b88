"""ORACLE (test infrastructure, not product): CPU float64 restatement of the reference's
discrete knowledge-gradient arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module.  The product path (``decoupled-kg_b200/``) never does.

Each function names the reference lines it follows (paths relative to
``/root/reference/src/decoupledbo/modules/acquisition/``):

===============================================  ==============================
oracle function                                  reference
===============================================  ==============================
``epigraph_indices``                             ``discretekg.py:341-412``
``expected_value_of_piecewise_linear_function``  ``discretekg.py:415-452``
``verify_intercepts_and_slopes``                 ``discretekg.py:455-470``
``kg_single_output`` (decoupled)                 ``discretekg.py:238-338``
``kg_coupled``                                   ``discretekg.py:162-235``
``forward``                                      ``discretekg.py:131-159``
``choose_best_objective``                        ``../acquisition_optimisation_strategy.py:143-163``
``make_std_grid``                                ``../utils.py:79-107``
===============================================  ==============================

Pinning status: ``epigraph_indices`` and ``expected_value_of_piecewise_linear_function`` are
pinned against (a) all 19 known-answer tests the reference holds for them
(``tests/modules/acquisition/test_discretekg.py:138-342``) and (b) golden vectors produced by
the reference's own two functions, imported in the build container with stubbed botorch names
(``oracle/make_golden.py`` -> ``tests/golden/epigraph_golden.npz``).  The GP-posterior stage is
a restatement of third-party code (see ``oracle/gp.py``).
"""

from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import gp as ogp

SLOPE_SHORTCUT_TOL = 1e-9  # discretekg.py:363


class OracleTensorDimensionError(Exception):
    """Stands in for ``botorch.exceptions.BotorchTensorDimensionError``."""


class OracleUnsupportedError(Exception):
    """Stands in for ``botorch.exceptions.UnsupportedError``."""


# ----------------------------------------------------------------------------------------------
# discretekg.py:455-470
# ----------------------------------------------------------------------------------------------
def verify_intercepts_and_slopes(intercepts: Tensor, slopes: Tensor) -> None:
    if intercepts.dim() != 1 or slopes.dim() != 1:
        raise OracleTensorDimensionError(
            f"Expected 'intercepts' and 'slopes' to both be one-dimensional tensors. "
            f"Got {intercepts.dim()=} and {slopes.dim()=}."
        )
    if intercepts.shape != slopes.shape:
        raise OracleTensorDimensionError(
            f"Expected 'intercepts' and 'slopes' to have the same shape. "
            f"Got {intercepts.shape=} and {slopes.shape=}."
        )
    if intercepts.shape[-1] == 0:
        raise ValueError(
            f"Expected inputs to specify at least one line. Got {intercepts.shape[-1]=}."
        )


# ----------------------------------------------------------------------------------------------
# discretekg.py:341-412  -- the march itself, on numpy float64 (no autograd needed for indices)
# ----------------------------------------------------------------------------------------------
def epigraph_march_np(a: np.ndarray, b: np.ndarray) -> Tuple[np.ndarray, bool]:
    """Indices (into the *original* arrays) of the lines on the upper envelope, left to right.

    Returns ``(indices, shortcut)``; ``shortcut`` is True when the all-slopes-tiny branch
    (discretekg.py:363-367) was taken, in which case ``indices == [argmax a]``.
    """
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if np.all(np.abs(b) < SLOPE_SHORTCUT_TOL):
        return np.array([int(np.argmax(a))], dtype=np.int64), True

    # slope ascending; ties by intercept descending (descending sort on a, then a *stable*
    # ascending sort on b -- discretekg.py:371-374)
    # torch's own sort primitives are used for the ordering so that index choices among
    # *identical* lines (same a and b) agree with the reference's unstable descending sort.
    ta = torch.from_numpy(a)
    tb = torch.from_numpy(b)
    first = torch.sort(ta, descending=True).indices
    second = torch.sort(tb[first], descending=False, stable=True).indices
    order = first[second].numpy()
    sa = a[order]
    sb = b[order]

    n = sa.shape[0]
    hull = [0]
    cur = 0
    while cur < n - 1:
        later = np.nonzero(sb[cur] != sb[cur + 1 :])[0]  # strictly different slope (:388)
        if later.size == 0:
            break
        cand = cur + 1 + later
        crossing = -(sa[cur] - sa[cand]) / (sb[cur] - sb[cand])  # (:395)
        pick = int(np.argmin(crossing))  # first minimum (:396)
        cur = int(cand[pick])
        hull.append(cur)
    return order[np.asarray(hull, dtype=np.int64)], False


def epigraph_indices(intercepts: Tensor, slopes: Tensor) -> Tuple[Tensor, Tensor]:
    """Upper envelope of the lines ``intercepts[n] + slopes[n] * z``.

    Returns ``(indices int64 [h], intersections [h-1])`` like the reference; ``intersections``
    is differentiable w.r.t. both inputs (same formula as discretekg.py:395 applied to the
    consecutive hull lines).  The empty intersections tensor is float64 (discretekg.py:366,404).
    """
    verify_intercepts_and_slopes(intercepts, slopes)
    idx_np, _shortcut = epigraph_march_np(
        intercepts.detach().cpu().numpy(), slopes.detach().cpu().numpy()
    )
    idx = torch.as_tensor(idx_np, dtype=torch.long, device=slopes.device)
    if idx.numel() <= 1:
        return idx, torch.tensor([], device=slopes.device, dtype=torch.double)
    lo, hi = idx[:-1], idx[1:]
    inter = -(intercepts[lo] - intercepts[hi]) / (slopes[lo] - slopes[hi])
    return idx, inter


# ----------------------------------------------------------------------------------------------
# discretekg.py:415-452
# ----------------------------------------------------------------------------------------------
_LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


def expected_value_of_piecewise_linear_function(
    intercepts: Tensor, slopes: Tensor, boundaries: Tensor
) -> Tensor:
    """E[f(Z)], Z ~ N(0,1), f piecewise linear with pieces ``intercepts[k] + slopes[k] z`` on
    ``(boundaries[k-1], boundaries[k])``."""
    verify_intercepts_and_slopes(intercepts, slopes)
    if boundaries.shape != (len(intercepts) - 1,):
        raise OracleTensorDimensionError(
            f"Expected 'boundaries' to be a one-dimensional tensor with "
            f"{len(intercepts)} elements. Got {boundaries.shape=}."
        )
    neg = torch.full((1,), -math.inf, dtype=boundaries.dtype, device=boundaries.device)
    pos = torch.full((1,), math.inf, dtype=boundaries.dtype, device=boundaries.device)
    z = torch.cat([neg, boundaries, pos])
    pdf = torch.exp(-0.5 * z * z - _LOG_SQRT_2PI)  # exp(Normal(0,1).log_prob(z))  (:442)
    cdf = 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))  # Normal(0,1).cdf(z)         (:443)
    return torch.sum(intercepts * (cdf[1:] - cdf[:-1]) - slopes * (pdf[1:] - pdf[:-1]))


def expected_max_of_lines(intercepts: Tensor, slopes: Tensor) -> Tensor:
    """E[max_n(a_n + b_n Z)]: hull then closed form (the body of the loop at :329-336)."""
    idx, inter = epigraph_indices(intercepts, slopes)
    return expected_value_of_piecewise_linear_function(intercepts[idx], slopes[idx], inter)


def expected_max_gradients_np(a: np.ndarray, b: np.ndarray):
    """Closed-form d E[max] / d a_n, d b_n on the hull (envelope theorem):
    ``dE/da_k = Phi(z_{k+1}) - Phi(z_k)``, ``dE/db_k = -(phi(z_{k+1}) - phi(z_k))``.
    Returns (E, idx, dE_da[h], dE_db[h], intersections[h-1])."""
    idx, shortcut = epigraph_march_np(a, b)
    ha, hb = a[idx], b[idx]
    if len(idx) > 1:
        x = -(ha[:-1] - ha[1:]) / (hb[:-1] - hb[1:])
    else:
        x = np.zeros(0)
    z = np.concatenate([[-np.inf], x, [np.inf]])
    pdf = np.exp(-0.5 * z * z - _LOG_SQRT_2PI)
    cdf = np.array([0.5 * (1.0 + math.erf(v / math.sqrt(2.0))) for v in z])
    dphi = cdf[1:] - cdf[:-1]
    dpdf = pdf[1:] - pdf[:-1]
    E = float(np.sum(ha * dphi - hb * dpdf))
    return E, idx, dphi, -dpdf, x


# ----------------------------------------------------------------------------------------------
# discretekg.py:238-338 (decoupled)
# ----------------------------------------------------------------------------------------------
def _check_weights(scalarisation_weights: Tensor) -> None:
    if scalarisation_weights.dim() != 2:
        raise OracleTensorDimensionError(
            "Expected 'scalarisation_weights' to have two dimensions: The first "
            "indexing different scalarisations to be averaged over and the second "
            "indexing coordinates of the objective space."
        )


def lines_single_output(
    model: ogp.OracleModelList,
    xnew: Tensor,
    obj_idx_new: int,
    discretisation: Tensor,
    scalarisation_weights: Tensor,
    dense: bool = True,
):
    """Intercepts and slopes ``(S, N+1)`` of discretekg.py:300-321; line 0 is xnew itself."""
    xnew = torch.as_tensor(xnew, dtype=torch.double)
    discretisation = torch.as_tensor(discretisation, dtype=torch.double)
    pts = torch.cat([xnew.unsqueeze(0), discretisation])
    means = []
    cov_i = var_i = None
    for m, obj in enumerate(model.models):
        if dense:
            mean_m, cov_m = ogp.posterior(obj, pts, observation_noise=False)
            if m == obj_idx_new:
                cov_i = cov_m[0]
                _, noisy = ogp.posterior(obj, xnew.unsqueeze(0), observation_noise=True)
                var_i = noisy[0, 0]
        else:
            mean_m, row, vn = ogp.posterior_row(obj, xnew, discretisation)
            if m == obj_idx_new:
                cov_i, var_i = row, vn
        means.append(mean_m.unsqueeze(-1))
    means = torch.cat(means, dim=-1)  # (N+1, M)
    znew = cov_i / var_i.sqrt()  # (:313)
    M = scalarisation_weights.shape[-1]
    w = scalarisation_weights.reshape(-1, 1, M)
    intercepts = torch.sum(w * means, dim=-1)  # (:320)
    slopes = w[..., obj_idx_new] * znew  # (:321)
    return intercepts, slopes


def kg_single_output(
    model: ogp.OracleModelList,
    xnew: Tensor,
    obj_idx_new: int,
    discretisation: Tensor,
    scalarisation_weights: Tensor,
    dense: bool = True,
) -> Tensor:
    _check_weights(scalarisation_weights)
    if not isinstance(model, ogp.OracleModelList):
        raise OracleUnsupportedError(f"Input 'model' must be a model list. Got {type(model)=}.")
    intercepts, slopes = lines_single_output(
        model, xnew, obj_idx_new, discretisation, scalarisation_weights, dense=dense
    )
    S = scalarisation_weights.shape[0]
    vals = []
    for j in range(S):
        e = expected_max_of_lines(intercepts[j], slopes[j])
        vals.append(e - torch.max(intercepts[j]))  # (:336)
    return torch.stack(vals).mean()  # (:338)


# ----------------------------------------------------------------------------------------------
# discretekg.py:162-235 (coupled)
# ----------------------------------------------------------------------------------------------
def lines_coupled(
    model: ogp.OracleModelList,
    xnew: Tensor,
    discretisation: Tensor,
    scalarisation_weights: Tensor,
    dense: bool = True,
):
    """Per-scalarisation intercepts/slopes ``(S, N+1)`` of discretekg.py:200-223.

    ``ModelListGP.posterior`` of independent sub-models is block diagonal over outputs, so
    ``ScalarizedPosteriorTransform(w)`` gives mean ``sum_m w_m mu_m`` and covariance
    ``sum_m w_m^2 Cov_m`` [BoTorch, recalled]; the noisy variance sums ``w_m^2 (var_m + noise_m)``.
    """
    xnew = torch.as_tensor(xnew, dtype=torch.double)
    discretisation = torch.as_tensor(discretisation, dtype=torch.double)
    pts = torch.cat([xnew.unsqueeze(0), discretisation])
    means, rows, noisy = [], [], []
    for obj in model.models:
        if dense:
            mean_m, cov_m = ogp.posterior(obj, pts, observation_noise=False)
            _, nz = ogp.posterior(obj, xnew.unsqueeze(0), observation_noise=True)
            rows.append(cov_m[0])
            noisy.append(nz[0, 0])
        else:
            mean_m, row, vn = ogp.posterior_row(obj, xnew, discretisation)
            rows.append(row)
            noisy.append(vn)
        means.append(mean_m)
    means = torch.stack(means, dim=-1)  # (N+1, M)
    rows = torch.stack(rows, dim=-1)  # (N+1, M)
    noisy = torch.stack(noisy)  # (M,)
    W = scalarisation_weights
    intercepts = means @ W.T  # (N+1, S)
    cov = rows @ (W**2).T
    var = (W**2) @ noisy  # (S,)
    slopes = cov / var.sqrt()
    return intercepts.T.contiguous(), slopes.T.contiguous()


def kg_coupled(
    model: ogp.OracleModelList,
    xnew: Tensor,
    discretisation: Tensor,
    scalarisation_weights: Tensor,
    dense: bool = True,
) -> Tensor:
    _check_weights(scalarisation_weights)
    intercepts, slopes = lines_coupled(
        model, xnew, discretisation, scalarisation_weights, dense=dense
    )
    vals = []
    for j in range(scalarisation_weights.shape[0]):
        e = expected_max_of_lines(intercepts[j], slopes[j])
        vals.append(e - torch.max(intercepts[j]))  # (:233)
    return torch.stack(vals).mean()  # (:235)


# ----------------------------------------------------------------------------------------------
# discretekg.py:131-159
# ----------------------------------------------------------------------------------------------
def forward(
    model: ogp.OracleModelList,
    X: Tensor,
    discretisation: Tensor,
    scalarisation_weights: Tensor,
    target_output_ix: Optional[int],
    dense: bool = True,
) -> Tensor:
    """``DiscreteKnowledgeGradient.forward``: X is ``(*b, 1, d)`` (or ``(b, d)`` rows) -> ``(*b)``."""
    X = torch.as_tensor(X, dtype=torch.double)
    if X.dim() == 2:
        X = X.unsqueeze(-2)
    assert X.shape[-2] == 1, "expected q=1"
    batch_shape, d = X.shape[:-2], X.shape[-1]
    if d != discretisation.shape[-1]:
        raise RuntimeError(
            f"Expected X to have last dimension matching 'self.x_discretisation'. "
            f"Got {X.shape[-1]=}, {discretisation.shape[-1]=}."
        )
    out = []
    for xnew in X.reshape(-1, d):  # the reference's Python loop (:145)
        if target_output_ix is not None:
            out.append(
                kg_single_output(
                    model, xnew, target_output_ix, discretisation, scalarisation_weights, dense
                )
            )
        else:
            out.append(kg_coupled(model, xnew, discretisation, scalarisation_weights, dense))
    return torch.stack(out).reshape(batch_shape)


# ----------------------------------------------------------------------------------------------
# acquisition_optimisation_strategy.py:143-163
# ----------------------------------------------------------------------------------------------
def choose_best_objective(candidates: Sequence[Tuple[int, Tensor, Tensor]], costs):
    best_i, best_x, best_v = max(
        candidates, key=lambda t: (max(float(t[-1]), 0.0) / float(costs[t[0]]), -float(costs[t[0]]))
    )
    return best_i, best_x, best_v / costs[best_i]


# ----------------------------------------------------------------------------------------------
# utils.py:79-107
# ----------------------------------------------------------------------------------------------
def make_std_grid(n_points_per_axis: int, n_dimensions: int) -> Tensor:
    """``npa**d x d`` grid on [0,1]^d, first coordinate slowest."""
    if n_dimensions <= 0:
        raise ValueError(f"Expected n_dimensions >= 1. Got {n_dimensions}.")
    axis = torch.linspace(0, 1, n_points_per_axis, dtype=torch.double)
    grids = torch.meshgrid(*([axis] * n_dimensions), indexing="ij")
    return torch.stack([g.reshape(-1) for g in grids], dim=-1)
