"""Spec layer that drives the hot path -- mirror of ``DiscreteKgOptimisationSpec`` in the
reference's ``src/decoupledbo/modules/acquisition_optimisation_strategy.py:166-273`` (same
constructor, same two methods, same return values and tie-breaking).

Only what the discrete-KG path uses is here; the HVKG / JES specs of the reference wrap other
BoTorch acquisition functions and are out of scope.

The one intended difference: because ``DiscreteKnowledgeGradient.forward`` now evaluates a whole
t-batch per launch, ``batch_limit`` may be set to ``num_restarts`` (the reference pins it to 1,
``bo_loop.py:127-129``, because its forward loops in Python) and the raw samples are evaluated
in one call (``init_batch_limit = raw_samples``) unless the caller overrides it.
"""

from __future__ import annotations

import logging
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from ..optim import optimize_acqf
from .acquisition.discretekg import DiscreteKnowledgeGradient
from .utils import make_torch_std_grid

logger = logging.getLogger(__name__)

TKWARGS = {"dtype": torch.double, "device": "cpu"}  # reference: pipeline/constants.py:8


def _get_standard_bounds(input_dim: int) -> Tensor:  # strategy.py:555-558
    return torch.tensor([[0.0] * input_dim, [1.0] * input_dim], **TKWARGS)


def choose_best_objective(candidates, costs):
    """``AcquisitionOptimisationSpec._choose_best_objective`` (strategy.py:143-163): maximise
    ``max(value, 0) / cost``; ties go to the cheaper objective, then to the first in order
    (Python ``max`` keeps the first maximal element).  Returns ``(i, x, value / cost)``."""
    best_i, best_x, best_acq_value = max(
        candidates, key=lambda x: (max(x[-1], 0) / costs[x[0]], -costs[x[0]])
    )
    best_acq_value_per_cost = best_acq_value / costs[best_i]
    return best_i, best_x, best_acq_value_per_cost


class DiscreteKgOptimisationSpec:
    def __init__(
        self,
        n_discretisation_points_per_axis: int,
        num_restarts: int,
        raw_samples: int,
        batch_limit: int,
        max_iter: int,
        init_batch_limit: Optional[int] = None,
    ):
        """Same arguments as reference strategy.py:167-194 (+ optional ``init_batch_limit``)."""
        self.n_discretisation_points_per_axis = n_discretisation_points_per_axis
        self.num_restarts = num_restarts
        self.raw_samples = raw_samples
        self.batch_limit = batch_limit
        self.max_iter = max_iter
        self.init_batch_limit = init_batch_limit

    _choose_best_objective = staticmethod(choose_best_objective)

    def _options(self):
        opts = {"batch_limit": self.batch_limit, "maxiter": self.max_iter}
        if self.init_batch_limit is not None:
            opts["init_batch_limit"] = self.init_batch_limit
        return opts

    def _discretisation(self, input_dim: int) -> Tensor:
        return make_torch_std_grid(self.n_discretisation_points_per_axis, input_dim, TKWARGS)

    def optimize_for_single_objective(
        self,
        model,
        costs: Union[Tensor, List],
        input_dim: int,
        *,
        scalarisation_weights: Tensor,
        **_unused_kwargs,
    ) -> Tuple[Tensor, int, Tensor]:
        """strategy.py:196-240: one acquisition function per objective, optimise each, then pick
        the objective with the best value per cost."""
        standard_bounds = _get_standard_bounds(input_dim)
        candidates = []
        for i in range(model.num_outputs):
            acq_func = DiscreteKnowledgeGradient(
                model,
                x_discretisation=self._discretisation(input_dim),
                scalarisation_weights=scalarisation_weights,
                target_output_ix=i,
            )
            candidate_x, acq_value = optimize_acqf(
                acq_function=acq_func,
                bounds=standard_bounds,
                q=1,
                num_restarts=self.num_restarts,
                raw_samples=self.raw_samples,
                options=self._options(),
            )
            if acq_value < 0:
                logger.warning(
                    "Optimal acquisition function value is negative: obj_index=%i, acq_value=%f",
                    i, acq_value,
                )
            candidates.append((i, candidate_x.detach(), acq_value.detach()))
            acq_func.invalidate()
        best_i, best_x, best_kg_per_cost = self._choose_best_objective(candidates, costs)
        return best_x, best_i, best_kg_per_cost

    def optimize_for_full_evaluation(
        self,
        model,
        input_dim: int,
        *,
        scalarisation_weights: Tensor,
        **_unused_kwargs,
    ) -> Tuple[Tensor, Tensor]:
        """strategy.py:242-273 (coupled evaluation)."""
        standard_bounds = _get_standard_bounds(input_dim)
        acq_func = DiscreteKnowledgeGradient(
            model,
            x_discretisation=self._discretisation(input_dim),
            scalarisation_weights=scalarisation_weights,
        )
        candidate_x, acq_value = optimize_acqf(
            acq_function=acq_func,
            bounds=standard_bounds,
            q=1,
            num_restarts=self.num_restarts,
            raw_samples=self.raw_samples,
            options=self._options(),
        )
        if acq_value < 0:
            logger.warning("Optimal acquisition function value is negative: acq_value=%f", acq_value)
        return candidate_x.detach(), acq_value.detach()
