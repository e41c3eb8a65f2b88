"""Spec layer that drives the hot path -- mirror of ``DiscreteKgOptimisationSpec`` in the
reference's ``src/decoupledbo/modules/acquisition_optimisation_strategy.py:166-273`` (same
constructor, same two methods, same return values and tie-breaking).

Only what the discrete-KG path uses is here; the HVKG / JES specs of the reference wrap other
BoTorch acquisition functions and are out of scope.

Intended differences (none changes a signature or a result):

* because ``DiscreteKnowledgeGradient.forward`` now evaluates a whole t-batch per launch,
  ``batch_limit`` may be set to ``num_restarts`` (the reference pins it to 1, ``bo_loop.py:127-129``,
  because its forward loops in Python) and the raw samples are evaluated in one call
  (``init_batch_limit = raw_samples``) unless the caller overrides it;
* ``concurrent_objectives`` (attribute, default True): the per-objective optimisations of
  ``optimize_for_single_objective`` (the loop at strategy.py:208) run side by side, one host thread and
  CUDA stream per objective, so the kernels of one objective fill the launch gaps of the other.  The
  starting points are still generated objective by objective, in order, so the global RNG is consumed
  exactly as in the serial loop and the chosen ``(x, objective, value)`` is identical;
* ``shard_group`` (attribute, default None / ``DKG_SHARD=1``): every rank of the process group runs
  this same code (SPMD) and each large t-batch is split across the ranks inside ``forward``; L-BFGS
  then sees identical ``(f, g)`` on every rank, so the iterates stay in lock-step without any
  broadcast.  Objectives run one after the other in that mode (one collective
  stream).
"""

from __future__ import annotations

import logging
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from ..optim import gen_batch_initial_conditions, optimize_acqf
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
        self.concurrent_objectives = True
        self.shard_group = None  # None: DiscreteKnowledgeGradient's default (DKG_SHARD=1 -> WORLD)

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
        acq_funcs = []
        for i in range(model.num_outputs):
            acq_func = DiscreteKnowledgeGradient(
                model,
                x_discretisation=self._discretisation(input_dim),
                scalarisation_weights=scalarisation_weights,
                target_output_ix=i,
            )
            if self.shard_group is not None:
                acq_func.shard_group = self.shard_group
            acq_funcs.append(acq_func)

        def optimise(i, initial_conditions=None):
            return optimize_acqf(
                acq_function=acq_funcs[i],
                bounds=standard_bounds,
                q=1,
                num_restarts=self.num_restarts,
                raw_samples=self.raw_samples,
                options=self._options(),
                batch_initial_conditions=initial_conditions,
            )

        sharded = any(a.shard_group is not None for a in acq_funcs)
        if self.concurrent_objectives and len(acq_funcs) > 1 and not sharded:
            results = self._optimise_concurrently(acq_funcs, standard_bounds, optimise)
        else:
            results = [optimise(i) for i in range(len(acq_funcs))]
        candidates = []
        for i, (candidate_x, acq_value) in enumerate(results):
            if acq_value < 0:
                logger.warning(
                    "Optimal acquisition function value is negative: obj_index=%i, acq_value=%f",
                    i, acq_value,
                )
            candidates.append((i, candidate_x.detach(), acq_value.detach()))
            acq_funcs[i].invalidate()
        best_i, best_x, best_kg_per_cost = self._choose_best_objective(candidates, costs)
        return best_x, best_i, best_kg_per_cost

    def _optimise_concurrently(self, acq_funcs, bounds, optimise):
        """Starting points objective by objective (RNG order of the serial loop), then one host thread
        and CUDA stream per objective for the L-BFGS phase."""
        from concurrent.futures import ThreadPoolExecutor

        inits = [
            gen_batch_initial_conditions(
                acq_function=a, bounds=bounds, q=1, num_restarts=self.num_restarts,
                raw_samples=self.raw_samples, options=self._options(),
            )
            for a in acq_funcs
        ]
        use_cuda = torch.cuda.is_available()
        dev = torch.cuda.current_device() if use_cuda else None

        def work(i):
            if not use_cuda:
                return optimise(i, inits[i])
            torch.cuda.set_device(dev)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                return optimise(i, inits[i])

        with ThreadPoolExecutor(max_workers=len(acq_funcs)) as pool:
            return list(pool.map(work, range(len(acq_funcs))))

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
