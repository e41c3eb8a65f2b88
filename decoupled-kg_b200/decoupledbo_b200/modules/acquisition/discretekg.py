"""Discrete multi-objective knowledge gradient -- B200-native drop-in.

Mirror of the reference module ``src/decoupledbo/modules/acquisition/discretekg.py``: the same
class, constructor, ``forward(X)`` t-batch semantics, module-level functions, errors and autograd
behaviour, with the arithmetic running as hand-written sm_100a CUDA behind the C-ABI of
``libdkg_b200.so`` (``include/dkg_b200.h``).  There is no CPU fallback: without the library and a
CUDA device the calls raise.

What changes relative to the reference (by design):

* ``forward`` evaluates ALL rows of the t-batch in one launch sequence instead of a Python loop
  over rows (``discretekg.py:145``); ``batch_limit``/``init_batch_limit`` of ``optimize_acqf`` can
  therefore be raised to the number of restarts / raw samples.
* everything that does not depend on the candidate (training Cholesky, posterior means at the
  discretisation, the scalarised intercept table, K^-1 k(X_train, X_disc)) is computed ONCE and
  cached ("plan"); the reference recomputes it per candidate.  The cache is keyed on a fingerprint
  of the model state and rebuilt when the model changes, so results always reflect the current
  model as they do in the reference.
* the (N+1) x (N+1) posterior covariance (``discretekg.py:301``) is never formed: only its first
  row is computed, as a tensor-core contraction.
* the backward is a fused closed-form (envelope theorem) kernel, not autograd through the graph.
* with ``shard_group`` set (or ``DKG_SHARD=1`` under ``torch.distributed``) the rows of a large
  t-batch are split across the ranks of the group and all-gathered on the device.

Both evaluation modes of the reference are covered: ``target_output_ix=i`` (decoupled,
``calculate_discrete_kg_conditioning_on_single_output``, ``discretekg.py:238-338``) and
``target_output_ix=None`` (coupled, ``calculate_discrete_kg``, ``discretekg.py:162-235``).
"""

from __future__ import annotations

import logging
import os
from collections import OrderedDict
from typing import Optional

import torch
from torch import Tensor

from ...botorch_compat import (
    AcquisitionFunction,
    BotorchTensorDimensionError,
    UnsupportedError,
    draw_sobol_samples,
    t_batch_mode_transform,
)
from ...gp_state import GPModelList, extract_gp_state, model_fingerprint
from ... import _native
from ... import distributed as _dist

logger = logging.getLogger(__name__)


def _raise_if_truncated(plan: "_native.Plan") -> None:
    lost = plan.stats()[6]
    raise RuntimeError(
        f"{lost} (candidate, scalarisation) sets have more upper-envelope vertices than the hull-record "
        f"spill pool holds; their gradient rows are NaN.  Raise DKG_SPILL_BLOCKS (32 vertices per block)."
    )


class _KGFunction(torch.autograd.Function):
    """kg[c] = KG(X[c]) with the gradient produced by the fused backward kernel."""

    @staticmethod
    def forward(ctx, X_flat: Tensor, plan: "_native.Plan", shard_group) -> Tensor:
        need_grad = bool(ctx.needs_input_grad[0])
        Xd = X_flat.detach()
        if shard_group is not None and _dist.should_shard(Xd.shape[0], shard_group):
            kg, dX = _dist.sharded_forward(plan, Xd, need_grad, shard_group)
        elif Xd.is_cuda:
            kg, dX = plan.forward_device(Xd, need_grad)
        else:
            kg, dX = plan.forward_host(Xd, need_grad)
        ctx.has_grad = need_grad
        if need_grad:
            # Sets with more than 64 hull vertices keep their extra records in a spill pool that grows on
            # demand; only if its size is pinned (DKG_SPILL_BLOCKS) can records be lost, and then the kernel
            # writes NaN into the gradient rows concerned.  The host entry point raises by itself (it
            # synchronises anyway); device results carry the NaN to the caller without forcing a sync
            # (DKG_CHECK=1 checks them here as well).
            if dX.is_cuda and os.environ.get("DKG_CHECK") == "1":
                if bool(torch.isnan(dX).any()) and not bool(torch.isnan(Xd).any()):
                    _raise_if_truncated(plan)
            ctx.save_for_backward(dX)
        return kg

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        if not ctx.has_grad:
            return None, None, None
        (dX,) = ctx.saved_tensors
        return grad_out.unsqueeze(-1) * dX, None, None


def _default_shard_group():
    if os.environ.get("DKG_SHARD") == "1" and _dist.is_active():
        return _dist.world_group()
    return None


class DiscreteKnowledgeGradient(AcquisitionFunction):
    """
    Discrete knowledge gradient (reference ``discretekg.py:25-159``).

    MOKG(x, d) = E[ max_x' E[f(x') . w | w, f_d(x') + eps] ] - E[ max_x' E[f(x') . w | w] ]
    """

    @classmethod
    def create_with_sobol_sample(
        cls,
        model,
        bounds: Tensor,
        num_discrete_points: int,
        scalarisation_weights: Optional[Tensor] = None,
        target_output_ix: Optional[int] = None,
    ):
        """Same contract as reference ``discretekg.py:33-60``: scrambled-Sobol discretisation."""
        x_discretisation = draw_sobol_samples(bounds, num_discrete_points, q=1)
        x_discretisation = x_discretisation.squeeze(1)
        x_discretisation = x_discretisation.to(bounds)
        return cls(model, x_discretisation, scalarisation_weights, target_output_ix)

    def __init__(
        self,
        model,
        x_discretisation: Tensor,
        scalarisation_weights: Optional[Tensor] = None,
        target_output_ix: Optional[int] = None,
    ):
        """
        Args (identical to reference ``discretekg.py:62-90``):
            model: A fitted ``ModelListGP`` (real BoTorch or duck-typed) or a ``GPModelList``.
            x_discretisation: ``k x d`` design points approximating the input space.
            scalarisation_weights: ``nscalar x n_objectives`` linear scalarisation weights
                (required for multi-output models).
            target_output_ix: if given, the KG assumes only this output is observed (decoupled).
        """
        super().__init__(model=model)

        if x_discretisation.dim() != 2:
            raise BotorchTensorDimensionError(
                f"Expected 'x_discretisation' to have two dimensions. "
                f"Got {x_discretisation.dim()=}."
            )

        if scalarisation_weights is None:
            if model.num_outputs != 1:
                raise UnsupportedError(
                    "Models with more than one output must specify 'scalarisation_weights'."
                )
            else:
                scalarisation_weights = torch.tensor([[1.0]]).to(x_discretisation)

        if scalarisation_weights.dim() != 2:
            raise BotorchTensorDimensionError(
                f"Expected 'scalarisation_weights' to have two dimensions: The first "
                f"indexing different scalarisations to be averaged over and the second "
                f"indexing coordinates of the objective space. "
                f"Got {scalarisation_weights.dim()=}"
            )
        if scalarisation_weights.shape[-1] != model.num_outputs:
            raise BotorchTensorDimensionError(
                f"Expected the last dimension of 'scalarisation_weights' to have one "
                f"element per objective. Got {scalarisation_weights.shape[-1]=} != "
                f"{model.num_outputs}=model.num_outputs."
            )

        self.x_discretisation = x_discretisation
        self.scalarisation_weights = scalarisation_weights
        self.target_output_ix = target_output_ix
        self._plan = None
        self._plan_key = None
        self._precision = "float64"
        # torch.distributed process group over whose ranks large t-batches are sharded (None: never)
        self.shard_group = _default_shard_group()

    def set_X_pending(self, X_pending: Optional[Tensor] = None) -> None:
        raise UnsupportedError(f"{type(self).__name__} does not account for X_pending yet.")

    # -- precision (an addition to the reference surface; the constructor is unchanged) ---------
    @property
    def precision(self) -> str:
        """``"float64"`` (default; rel 1e-9 against the reference) or ``"float32"``: the covariance
        contraction keeps 4 base-256 digits per operand (10 int8 digit products instead of 34), the
        rest of the path stays float64.  Tolerance in that mode: ``|dKG| <= 1e-4 |KG| + 1e-7 max|a|``."""
        return self._precision

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ("float64", "float32"):
            raise ValueError(f"precision must be 'float64' or 'float32'; got {value!r}")
        if value != self._precision:
            self.invalidate()
            self._precision = value

    # -- native state -------------------------------------------------------------------------
    def _state_key(self):
        xd, W = self.x_discretisation, self.scalarisation_weights
        return (
            model_fingerprint(self.model),
            (xd.data_ptr(), xd._version, tuple(xd.shape)),
            (W.data_ptr(), W._version, tuple(W.shape)),
            self.target_output_ix,
            self._precision,
        )

    def _get_plan(self) -> "_native.Plan":
        """The candidate-independent GPU state for this acquisition function: built at first use and
        rebuilt whenever the model, the discretisation or the weights changed since."""
        if not hasattr(self.model, "models"):
            raise UnsupportedError(
                f"Input 'model' must be a 'ModelListGP'. Got {type(self.model)=}."
            )
        key = self._state_key()
        if self._plan is None or key != self._plan_key:
            self.invalidate()
            self._plan = _build_plan(
                self.model, self.x_discretisation, self.scalarisation_weights, self.target_output_ix,
                self._precision,
            )
            self._plan_key = key
        return self._plan

    def append_observation(self, objective_ix: int, x: Tensor, y, model=None) -> bool:
        """Incremental refresh between BO iterations (an addition to the reference surface): objective
        ``objective_ix`` received the observation ``(x, y)`` and the hyper-parameters did not change
        (``--fit-hyperparams=never`` / ``once``: ``bo_loop.py:403-405``, ``:574-589``).  The cached GPU state
        is extended in O(n^2 + n N) instead of being rebuilt in O(n^3 + n^2 N).

        ``y`` is in the space of the model's ``train_targets`` (standardised if an outcome transform is
        used).  ``model``: the model that now includes the point -- it replaces ``self.model``; if omitted
        and ``self.model`` is a ``GPModelList`` the point is appended to it here.  Returns True when the
        state was extended in place, False when it has to be rebuilt at the next call (no plan yet, no
        room left, or the extended covariance is not positive definite without new jitter)."""
        x = torch.as_tensor(x, dtype=torch.double).detach().reshape(-1)
        yv = float(torch.as_tensor(y).reshape(-1)[0])
        if model is not None:
            if isinstance(model, torch.nn.Module):
                self.model = model
            else:  # tensor-only GP state containers are plain objects (see botorch_compat.AcquisitionFunction)
                object.__setattr__(self, "model", model)
        elif isinstance(self.model, GPModelList):
            o = self.model.models[objective_ix]
            o.train_x = torch.cat([o.train_x, x.reshape(1, -1).to(o.train_x)])
            o.train_y = torch.cat([o.train_y, torch.tensor([yv], dtype=o.train_y.dtype)])
        if self._plan is None:
            return False
        try:
            self._plan.append_point(objective_ix, x, yv)
        except (_native.PlanCapacityError, RuntimeError):
            self.invalidate()
            return False
        self._plan_key = self._state_key()  # the model now matches the extended plan
        return True

    def invalidate(self) -> None:
        """Drop the cached GPU state (done automatically when the model state changes)."""
        if self._plan is not None:
            self._plan.close()
        self._plan = None
        self._plan_key = None

    @t_batch_mode_transform(expected_q=1)
    def forward(self, X: Tensor) -> Tensor:
        """``X``: ``(*b) x 1 x d`` -> KG values of shape ``(*b)`` (reference ``:131-159``)."""
        batch_shape, d = X.shape[:-2], X.shape[-1]

        if d != self.x_discretisation.shape[-1]:
            raise RuntimeError(
                f"Expected X to have last dimension matching 'self.x_discretisation'. "
                f"Got {X.shape[-1]=}, {self.x_discretisation.shape[-1]=}."
            )
        X_flat = X.reshape(-1, d)
        if X_flat.dtype != torch.double:
            X_flat = X_flat.to(torch.double)
        kgvals = _KGFunction.apply(X_flat, self._get_plan(), self.shard_group)
        return kgvals.to(X.dtype).reshape(batch_shape)


def _build_plan(model, x_discretisation, scalarisation_weights, target_output_ix, precision="float64"):
    state: GPModelList = extract_gp_state(model)
    if target_output_ix is not None:
        # the reference indexes Python lists / tensors with it (discretekg.py:301, :321), so negative
        # values count from the end and anything outside [-M, M) is an IndexError
        M = state.num_outputs
        ix = int(target_output_ix)
        if not -M <= ix < M:
            raise IndexError(f"target_output_ix={target_output_ix} is out of range for {M} objectives")
        target_output_ix = ix % M
    flags = _native.PLAN_FAST32 if precision == "float32" else _native.PLAN_DEFAULT
    return _native.Plan(state, x_discretisation, scalarisation_weights, target_output_ix, flags)


# ---------------------------------------------------------------------------------------------
# Module-level functions of the reference (``discretekg.py:162-470``), GPU-backed.  The reference's
# own test-suite imports them by name (``tests/modules/acquisition/test_discretekg.py:7-13``).
# ---------------------------------------------------------------------------------------------
_PLAN_CACHE: "OrderedDict[tuple, _native.Plan]" = OrderedDict()
_PLAN_CACHE_SIZE = 4


def _cached_plan(model, discretisation, scalarisation_weights, target_output_ix):
    key = (
        model_fingerprint(model),
        (discretisation.data_ptr(), discretisation._version, tuple(discretisation.shape)),
        (scalarisation_weights.data_ptr(), scalarisation_weights._version, tuple(scalarisation_weights.shape)),
        target_output_ix,
    )
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = _build_plan(model, discretisation, scalarisation_weights, target_output_ix)
        # the key holds storage pointers: keep the tensors alive for as long as the entry lives
        plan._cache_refs = (model, discretisation, scalarisation_weights)
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _, old = _PLAN_CACHE.popitem(last=False)
            old.close()
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def _check_weights(scalarisation_weights):
    if scalarisation_weights.dim() != 2:
        raise BotorchTensorDimensionError(
            "Expected 'scalarisation_weights' to have two dimensions: The first "
            "indexing different scalarisations to be averaged over and the second "
            "indexing coordinates of the objective space."
        )


def _kg_at_point(model, xnew, discretisation, scalarisation_weights, target_output_ix):
    plan = _cached_plan(model, discretisation, scalarisation_weights, target_output_ix)
    X_flat = xnew.reshape(1, -1)
    if X_flat.dtype != torch.double:
        X_flat = X_flat.to(torch.double)
    kg = _KGFunction.apply(X_flat, plan, None)
    return kg[0].to(scalarisation_weights.dtype)


def calculate_discrete_kg(model, xnew, discretisation, scalarisation_weights):
    """Coupled discrete KG at one point (reference ``discretekg.py:162-235``): ``xnew`` is a 1-D
    tensor, the result a scalar tensor differentiable with respect to ``xnew``."""
    _check_weights(scalarisation_weights)
    if not hasattr(model, "models"):
        # the reference accepts any model returning a GPyTorchPosterior here; the CUDA path needs the
        # per-objective exact-GP state, i.e. a model list
        raise UnsupportedError(f"Input 'model' must be a 'ModelListGP'. Got {type(model)=}.")
    return _kg_at_point(model, xnew, discretisation, scalarisation_weights, None)


def calculate_discrete_kg_conditioning_on_single_output(
    model, xnew, obj_idx_new, discretisation, scalarisation_weights
):
    """Decoupled discrete KG at one point when only objective ``obj_idx_new`` is observed
    (reference ``discretekg.py:238-338``)."""
    _check_weights(scalarisation_weights)
    if not hasattr(model, "models"):
        raise UnsupportedError(f"Input 'model' must be a 'ModelListGP'. Got {type(model)=}.")
    return _kg_at_point(model, xnew, discretisation, scalarisation_weights, int(obj_idx_new))


def calculate_epigraph_indices(intercepts: Tensor, slopes: Tensor):
    """Upper envelope ("epigraph") of the lines ``intercepts + slopes * z`` (reference
    ``discretekg.py:341-412``): ``(indices, intersections)``, left to right.

    The envelope is found on the GPU (``dkg_expected_max_lines_dev``: chord filter + the reference's
    march with the same ordering, strict-slope filter, division and tie-breaks).  The intersections
    are then re-formed from the input tensors with the reference's expression (``:395``), so they
    carry the same values and the same autograd graph as the reference's.
    """
    _verify_intercepts_and_slopes(intercepts, slopes)

    device = slopes.device
    cap = 64
    while True:
        res = _native.expected_max_lines(intercepts.unsqueeze(0), slopes.unsqueeze(0), hull_cap=cap)
        h = int(res["hull_count"][0])
        if h <= cap:
            break
        cap = h
    indices = res["hull_idx"][0, :h].to(device=device, dtype=torch.long)
    if h == 1:
        # (also the |slopes| < 1e-9 shortcut, :363-367)
        return indices, torch.tensor([], device=device, dtype=torch.double)
    i, j = indices[:-1], indices[1:]
    intersections = -(intercepts[i] - intercepts[j]) / (slopes[i] - slopes[j])
    return indices, intersections


class _PiecewiseExpectation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, intercepts: Tensor, slopes: Tensor, boundaries: Tensor) -> Tensor:
        want = any(ctx.needs_input_grad)
        res = _native.piecewise_expectation(
            intercepts.reshape(1, -1), slopes.reshape(1, -1), boundaries.reshape(1, -1), want_grad=want
        )
        ctx.has_grad = want
        if want:
            ctx.save_for_backward(
                res["dE_da"][0].to(intercepts), res["dE_db"][0].to(slopes), res["dE_dz"][0].to(boundaries)
            )
        return res["e"][0].to(device=boundaries.device, dtype=intercepts.dtype)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        if not ctx.has_grad:
            return None, None, None
        da, db, dz = ctx.saved_tensors
        need = ctx.needs_input_grad
        return (
            grad_out * da if need[0] else None,
            grad_out * db if need[1] else None,
            grad_out * dz if need[2] else None,
        )


def calculate_expected_value_of_piecewise_linear_function(
    intercepts: Tensor, slopes: Tensor, boundaries: Tensor
):
    """E[f(Z)], Z ~ N(0, 1), for the piecewise-linear f with the given pieces and break points
    (reference ``discretekg.py:415-452``); differentiable in all three arguments."""
    _verify_intercepts_and_slopes(intercepts, slopes)
    if boundaries.shape != (len(intercepts) - 1,):
        raise BotorchTensorDimensionError(
            f"Expected 'boundaries' to be a one-dimensional tensor with "
            f"{len(intercepts)} elements. Got {boundaries.shape=}."
        )
    if not intercepts.is_floating_point():
        intercepts = intercepts.to(torch.get_default_dtype())
    if not slopes.is_floating_point():
        slopes = slopes.to(torch.get_default_dtype())
    return _PiecewiseExpectation.apply(intercepts, slopes, boundaries)


def _verify_intercepts_and_slopes(intercepts, slopes):
    if intercepts.dim() != 1 or slopes.dim() != 1:
        raise BotorchTensorDimensionError(
            f"Expected 'intercepts' and 'slopes' to both be one-dimensional tensors. "
            f"Got {intercepts.dim()=} and {slopes.dim()=}."
        )
    if intercepts.shape != slopes.shape:
        raise BotorchTensorDimensionError(
            f"Expected 'intercepts' and 'slopes' to have the same shape. "
            f"Got {intercepts.shape=} and {slopes.shape=}."
        )
    if intercepts.shape[-1] == 0:
        raise ValueError(
            f"Expected inputs to specify at least one line. "
            f"Got {intercepts.shape[-1]=}."
        )
