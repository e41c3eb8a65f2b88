"""``optimize_acqf`` for the spec layer.

The reference calls ``botorch.optim.optimize_acqf`` (third-party; ``acquisition_optimisation_
strategy.py:217-224``).  When BoTorch is importable that function is used unchanged.  Otherwise
this module provides the same algorithm for the case the spec uses (``q=1``, box bounds, no
constraints) [BoTorch @ c14808f, recalled -- SURVEY.md Appendix A]:

1. ``raw_samples`` scrambled-Sobol points, acquisition values under ``no_grad`` in chunks of
   ``options["init_batch_limit"]`` (default: ``batch_limit``);
2. ``initialize_q_batch``: keep the best raw point, sample the other starts with probability
   proportional to ``exp(eta * zscore)``;
3. L-BFGS-B (scipy) on each chunk of ``batch_limit`` restarts, objective ``-acqf(X).sum()``,
   gradient from autograd (here: the fused CUDA backward), ``maxiter`` iterations;
4. final values under ``no_grad``; the best restart by first-index argmax.
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from .botorch_compat import HAVE_BOTORCH, draw_sobol_samples


def initialize_q_batch(X: Tensor, Y: Tensor, n: int, eta: float = 2.0) -> Tensor:
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError(f"n ({n}) cannot be larger than the number of provided samples ({n_samples})")
    if n == n_samples:
        return X
    Ystd = Y.std()
    if Ystd == 0 or not torch.isfinite(Ystd):
        return X[torch.randperm(n=n_samples, device=X.device)][:n]
    max_val, max_idx = torch.max(Y, dim=0)
    Z = (Y - Y.mean()) / Ystd
    etaZ = eta * Z
    weights = torch.exp(etaZ)
    while torch.isinf(weights).any():
        etaZ *= 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, n)
    if max_idx not in idcs:
        idcs[-1] = max_idx
    return X[idcs]


def _batched_values(acq_function, X: Tensor, limit: int) -> Tensor:
    out = []
    with torch.no_grad():
        for lo in range(0, X.shape[0], limit):
            out.append(acq_function(X[lo : lo + limit]))
    return torch.cat(out)


def _lbfgsb_chunk(acq_function, X0: Tensor, bounds: Tensor, maxiter: int) -> Tensor:
    from scipy.optimize import minimize

    shape = X0.shape
    lo = bounds[0].expand(shape).reshape(-1).cpu().numpy()
    hi = bounds[1].expand(shape).reshape(-1).cpu().numpy()

    def f_and_g(x_np: np.ndarray):
        X = torch.from_numpy(x_np.reshape(shape)).to(X0).contiguous().requires_grad_(True)
        loss = -acq_function(X).sum()
        (g,) = torch.autograd.grad(loss, X)
        return float(loss.detach()), g.reshape(-1).cpu().numpy().astype(np.float64)

    res = minimize(
        f_and_g, X0.reshape(-1).cpu().numpy().astype(np.float64), jac=True, method="L-BFGS-B",
        bounds=list(zip(lo, hi)), options={"maxiter": maxiter},
    )
    X = torch.from_numpy(res.x.reshape(shape)).to(X0)
    return torch.max(torch.min(X, bounds[1]), bounds[0])


def gen_batch_initial_conditions(
    acq_function, bounds: Tensor, q: int, num_restarts: int, raw_samples: int, options: Optional[Dict] = None
) -> Tensor:
    """Steps 1-2 above (BoTorch's function of the same name): ``num_restarts x q x d`` starting points.
    Draws from the global torch RNG (Sobol seed, multinomial), like BoTorch's."""
    if HAVE_BOTORCH:  # pragma: no cover
        from botorch.optim.initializers import gen_batch_initial_conditions as _botorch_gen

        return _botorch_gen(acq_function=acq_function, bounds=bounds, q=q, num_restarts=num_restarts,
                            raw_samples=raw_samples, options=options)
    if q != 1:
        raise NotImplementedError("only q=1 is used by DiscreteKgOptimisationSpec")
    options = dict(options or {})
    batch_limit = int(options.get("batch_limit", num_restarts))
    init_limit = int(options.get("init_batch_limit", batch_limit))
    seed = int(torch.randint(0, 2**31 - 1, (1,)).item())
    X_raw = draw_sobol_samples(bounds=bounds, n=raw_samples, q=q, seed=seed)  # raw x 1 x d
    Y_raw = _batched_values(acq_function, X_raw, init_limit)
    return initialize_q_batch(X_raw, Y_raw, n=num_restarts, eta=float(options.get("eta", 2.0)))


def optimize_acqf(
    acq_function,
    bounds: Tensor,
    q: int,
    num_restarts: int,
    raw_samples: Optional[int] = None,
    options: Optional[Dict] = None,
    batch_initial_conditions: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor]:
    """Same call as the reference makes (``q=1``); returns ``(candidate 1 x d, value)``.
    ``batch_initial_conditions`` (``num_restarts x 1 x d``) skips the raw-sample phase, as in BoTorch."""
    if HAVE_BOTORCH:  # pragma: no cover
        from botorch.optim import optimize_acqf as _botorch_optimize_acqf

        return _botorch_optimize_acqf(
            acq_function=acq_function, bounds=bounds, q=q, num_restarts=num_restarts,
            raw_samples=raw_samples, options=options, batch_initial_conditions=batch_initial_conditions,
        )
    if q != 1:
        raise NotImplementedError("only q=1 is used by DiscreteKgOptimisationSpec")
    options = dict(options or {})
    batch_limit = int(options.get("batch_limit", num_restarts))
    maxiter = int(options.get("maxiter", 200))
    X_init = batch_initial_conditions
    if X_init is None:
        X_init = gen_batch_initial_conditions(acq_function, bounds, q, num_restarts, raw_samples, options)
    chunks = []
    for lo in range(0, num_restarts, batch_limit):
        chunks.append(_lbfgsb_chunk(acq_function, X_init[lo : lo + batch_limit], bounds, maxiter))
    cands = torch.cat(chunks)
    vals = _batched_values(acq_function, cands, batch_limit)
    best = int(torch.argmax(vals))
    return cands[best], vals[best]
