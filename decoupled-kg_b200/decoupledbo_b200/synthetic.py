"""Seeded synthetic GP problems of the shapes BASELINE.json names.

Pure tensor generators (CPU, float64, deterministic): they produce *inputs* only — GP state,
discretisation, scalarisation weights, candidates — for the tests and for ``bench.py``.

* ``problem_c2``  d=2, n=100, N=1024 (32x32 grid), S=16, C=512.  GP state = the generator used
  by the reference notebook for the "lengthscales" family (``fixed_hyperparams`` of
  ``data/shared/gp-problem/lengthscales/*.pt``: Matern-5/2, lengthscales [0.2, 1.8], output
  scales [1, 50], zero mean) on seeded Sobol training inputs with a seeded GP-prior draw as
  targets.  (The committed ``.pt`` files live in /root/reference, which does not travel to the
  GPU box, hence the regenerated state.)
* ``problem_c4``  d=4, n=400, N=16384 (scrambled Sobol), S=16, C=4096; "observationnoise" family:
  lengthscale 0.4, output scale 1, observation noise sd [1, 0] -> model noise [1.0, 1e-4].
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch
from torch import Tensor

from .gp_state import MATERN52, GPObjective, GPModelList


def sobol(n: int, d: int, seed: int) -> Tensor:
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    return eng.draw(n, dtype=torch.double)


def simplex_weights(n_scal: int, n_obj: int = 2, seed: int = 0) -> Tensor:
    """What ``sample_simplex(n_obj, n_scal, qmc=True)`` yields (``bo_loop.py:98-116``): scrambled
    Sobol in ``n_obj-1`` dims, sorted, consecutive differences.  For two objectives: (u, 1-u)."""
    if n_obj == 1:
        return torch.ones(n_scal, 1, dtype=torch.double)
    u = sobol(n_scal, n_obj - 1, seed)
    u, _ = torch.sort(u, dim=-1)
    pad0 = torch.zeros(n_scal, 1, dtype=torch.double)
    pad1 = torch.ones(n_scal, 1, dtype=torch.double)
    edges = torch.cat([pad0, u, pad1], dim=-1)
    return edges[:, 1:] - edges[:, :-1]


def _matern52(x1: Tensor, x2: Tensor, ls: Tensor, s: float) -> Tensor:
    r = torch.cdist(x1 / ls, x2 / ls)
    return s * (1.0 + math.sqrt(5.0) * r + 5.0 / 3.0 * r * r) * torch.exp(-math.sqrt(5.0) * r)


def gp_prior_draw(x: Tensor, ls: Tensor, s: float, seed: int) -> Tensor:
    K = _matern52(x, x, ls, s) + 1e-10 * torch.eye(x.shape[0], dtype=torch.double)
    L = torch.linalg.cholesky(K)
    g = torch.Generator().manual_seed(seed)
    return L @ torch.randn(x.shape[0], generator=g, dtype=torch.double)


@dataclass
class SyntheticProblem:
    name: str
    model: GPModelList
    x_disc: Tensor  # (N, d)
    weights: Tensor  # (S, M)
    candidates: Tensor  # (C, d)

    @property
    def d(self) -> int:
        return self.x_disc.shape[1]


def make_problem(
    name: str,
    d: int,
    n_train: int,
    lengthscales: List[float],
    outputscales: List[float],
    obs_noise_sd: List[float],
    model_noise: List[float],
    x_disc: Tensor,
    n_scal: int,
    n_cand: int,
    seed_train: int,
    seed_cand: int,
    seed_w: int = 0,
) -> SyntheticProblem:
    xtr = sobol(n_train, d, seed_train)
    objs = []
    for m in range(len(lengthscales)):
        ls = torch.full((d,), lengthscales[m], dtype=torch.double)
        y = gp_prior_draw(xtr, ls, outputscales[m], seed_train + 1 + m)
        if obs_noise_sd[m] > 0:
            g = torch.Generator().manual_seed(seed_train + 101 + m)
            y = y + obs_noise_sd[m] * torch.randn(n_train, generator=g, dtype=torch.double)
        objs.append(
            GPObjective(
                train_x=xtr.clone(),
                train_y=y,
                lengthscale=ls,
                outputscale=outputscales[m],
                mean_const=0.0,
                noise=model_noise[m],
                kernel=MATERN52,
            )
        )
    return SyntheticProblem(
        name=name,
        model=GPModelList(objs),
        x_disc=x_disc,
        weights=simplex_weights(n_scal, len(objs), seed_w),
        candidates=sobol(n_cand, d, seed_cand),
    )


def std_grid(n_points_per_axis: int, d: int) -> Tensor:
    from .modules.utils import make_torch_std_grid

    return make_torch_std_grid(n_points_per_axis, d, {"dtype": torch.double})


def problem_c2(n_cand: int = 512, n_scal: int = 16, grid: int = 32, n_train: int = 100):
    """BASELINE.json configs[1]/[2]: 2-obj d=2 GP, n=100, |X_disc|=1024, 16 scalarisations."""
    return make_problem(
        "c2", 2, n_train, [0.2, 1.8], [1.0, 50.0], [0.0, 0.0], [1e-4, 1e-4],
        std_grid(grid, 2), n_scal, n_cand, seed_train=1234, seed_cand=1,
    )


def problem_c4(n_cand: int = 4096, n_scal: int = 16, n_disc: int = 16384, n_train: int = 400):
    """BASELINE.json configs[3]: 2-obj d=4, n=400, |X_disc|=16384, 16 scalarisations."""
    return make_problem(
        "c4", 4, n_train, [0.4, 0.4], [1.0, 1.0], [1.0, 0.0], [1.0, 1e-4],
        sobol(n_disc, 4, 7), n_scal, n_cand, seed_train=4242, seed_cand=8,
    )


def to_oracle_kwargs(model: GPModelList):
    """Plain dicts (tensors/floats) a test can feed to ``oracle.gp.OracleObjective``."""
    return [
        dict(
            train_x=o.train_x, train_y=o.train_y, lengthscale=o.lengthscale,
            outputscale=o.outputscale, mean_const=o.mean_const, noise=o.noise,
            kernel=o.kernel, y_mean=o.y_mean, y_std=o.y_std,
        )
        for o in model.models
    ]
