"""GP state the CUDA path consumes: plain tensors + hyper-parameters per objective.

The reference hands ``DiscreteKnowledgeGradient`` a BoTorch ``ModelListGP`` built by
``src/decoupledbo/modules/model/factory.py:24-135`` (per objective: ``SingleTaskGP`` on
unit-cube inputs, ``ConstantMean``, ``ScaleKernel(Matern-5/2 | RBF, ARD)``,
``GaussianLikelihood``, optional ``Standardize(m=1)``).  ``extract_gp_state`` reads exactly
those attributes from such a model (duck-typed, so it works with the real BoTorch classes when
they are importable); ``GPModelList`` is the tensor-only equivalent that can be passed as
``model`` directly when BoTorch is absent.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

import torch
from torch import Tensor

MATERN52 = 0
RBF = 1


@dataclass
class GPObjective:
    """One objective's exact-GP state (everything in *model* space)."""

    train_x: Tensor  # (n, d) unit cube
    train_y: Tensor  # (n,)   standardised if an outcome transform is used
    lengthscale: Tensor  # (d,)
    outputscale: float
    mean_const: float
    noise: float
    kernel: int = MATERN52
    y_mean: float = 0.0
    y_std: float = 1.0

    def __post_init__(self):
        self.train_x = torch.as_tensor(self.train_x, dtype=torch.double).detach()
        if self.train_x.dim() != 2:
            raise ValueError(f"train_x must be (n, d); got {tuple(self.train_x.shape)}")
        self.train_y = torch.as_tensor(self.train_y, dtype=torch.double).detach().reshape(-1)
        if self.train_y.numel() != self.train_x.shape[0]:
            raise ValueError("train_y must have one entry per training input")
        ls = torch.as_tensor(self.lengthscale, dtype=torch.double).detach().reshape(-1)
        if ls.numel() == 1:
            ls = ls.expand(self.train_x.shape[1]).clone()
        if ls.numel() != self.train_x.shape[1]:
            raise ValueError("lengthscale must be scalar or have one entry per input dimension")
        self.lengthscale = ls
        self.outputscale = float(self.outputscale)
        self.mean_const = float(self.mean_const)
        self.noise = float(self.noise)
        self.y_mean = float(self.y_mean)
        self.y_std = float(self.y_std)
        if self.kernel not in (MATERN52, RBF):
            raise ValueError(f"unsupported kernel id {self.kernel}")

    @property
    def n(self) -> int:
        return self.train_x.shape[0]

    @property
    def d(self) -> int:
        return self.train_x.shape[1]


class GPModelList:
    """Tensor-only stand-in for ``botorch.models.ModelListGP`` (``.models``, ``.num_outputs``)."""

    def __init__(self, models: Sequence[GPObjective]):
        self.models = list(models)
        if not self.models:
            raise ValueError("need at least one objective")
        d = self.models[0].d
        if any(o.d != d for o in self.models):
            raise ValueError("all objectives must share the input dimension")

    @property
    def num_outputs(self) -> int:
        return len(self.models)


def _scalar(x: Any) -> float:
    return float(torch.as_tensor(x).detach().reshape(-1)[0])


def _extract_single(gp: Any) -> GPObjective:
    """Read one (duck-typed) ``SingleTaskGP``.

    Supported surface (what ``factory.py:63-135`` builds): unbatched exact GP, ``ConstantMean``,
    ``ScaleKernel(Matern-5/2 | RBF)``, homoskedastic ``GaussianLikelihood``, optional
    ``Standardize(m=1)`` outcome transform, NO input transform.  Anything else raises
    ``NotImplementedError`` rather than being mis-read (the reference goes through
    ``model.posterior`` and is correct for any model).
    """
    train_x = gp.train_inputs[0]
    train_y = gp.train_targets
    if train_x.dim() != 2:
        raise NotImplementedError("batched sub-models are not supported")
    if getattr(gp, "input_transform", None) is not None:
        raise NotImplementedError("models with an input_transform are not supported (factory.py:65 normalises the data itself)")
    noise = torch.as_tensor(gp.likelihood.noise).detach()
    if noise.numel() != 1:
        raise NotImplementedError(
            f"only a homoskedastic GaussianLikelihood is supported; got a noise tensor of shape {tuple(noise.shape)} "
            f"(FixedNoise / heteroskedastic likelihoods are not)"
        )
    mean_module = gp.mean_module
    if not hasattr(mean_module, "constant") or torch.as_tensor(mean_module.constant).numel() != 1:
        raise NotImplementedError(f"only ConstantMean is supported (factory.py:72); got {type(mean_module).__name__}")
    covar = gp.covar_module
    base = getattr(covar, "base_kernel", None)
    if base is None:
        raise NotImplementedError("expected ScaleKernel(base_kernel) as built by factory.py:110-135")
    name = type(base).__name__.lower()
    if "matern" in name:
        nu = float(getattr(base, "nu", 2.5))
        if abs(nu - 2.5) > 1e-12:
            raise NotImplementedError(f"only Matern nu=2.5 is supported (factory config); got {nu}")
        kernel = MATERN52
    elif "rbf" in name:
        kernel = RBF
    else:
        raise NotImplementedError(f"unsupported base kernel {type(base).__name__}")
    ot = getattr(gp, "outcome_transform", None)
    y_mean, y_std = 0.0, 1.0
    if ot is not None:
        y_mean, y_std = _scalar(ot.means), _scalar(ot.stdvs)
    return GPObjective(
        train_x=train_x,
        train_y=train_y,
        lengthscale=base.lengthscale.detach().reshape(-1),
        outputscale=_scalar(covar.outputscale),
        mean_const=_scalar(gp.mean_module.constant),
        noise=_scalar(gp.likelihood.noise),
        kernel=kernel,
        y_mean=y_mean,
        y_std=y_std,
    )


def _fp_tensor(t: Tensor):
    return (t.data_ptr(), t._version, tuple(t.shape))


def _fp_objective(o: GPObjective):
    return (_fp_tensor(o.train_x), _fp_tensor(o.train_y), _fp_tensor(o.lengthscale), o.outputscale,
            o.mean_const, o.noise, o.kernel, o.y_mean, o.y_std)


def model_fingerprint(model: Any):
    """Cheap identity of everything ``extract_gp_state`` reads: storage pointers and in-place version
    counters of the training data and of every parameter / buffer (plain floats by value).  The
    acquisition function compares it on every call and rebuilds its cached GPU state when the model
    was re-fitted, re-conditioned in place or had hyper-parameters changed -- the reference queries
    the model on every call (``discretekg.py:275-284``), so a stale cache would be a silent divergence."""
    subs = getattr(model, "models", None)
    if subs is None:
        raise TypeError(f"expected a model list with a '.models' attribute; got {type(model)}")
    out = []
    for gp in subs:
        if isinstance(gp, GPObjective):
            out.append(_fp_objective(gp))
        elif isinstance(gp, torch.nn.Module):
            fp = [id(gp), _fp_tensor(gp.train_inputs[0]), _fp_tensor(gp.train_targets)]
            fp += [(t.data_ptr(), t._version) for t in gp.parameters()]
            fp += [(t.data_ptr(), t._version) for t in gp.buffers()]
            out.append(tuple(fp))
        else:  # duck-typed stand-in without parameters(): hyper-parameters by value
            o = _extract_single(gp)
            out.append((_fp_tensor(gp.train_inputs[0]), _fp_tensor(gp.train_targets), tuple(o.lengthscale.tolist()),
                        o.outputscale, o.mean_const, o.noise, o.kernel, o.y_mean, o.y_std))
    return tuple(out)


def extract_gp_state(model: Any) -> GPModelList:
    """``ModelListGP`` (real BoTorch or duck-typed) or ``GPModelList`` -> ``GPModelList``."""
    if isinstance(model, GPModelList):
        return model
    subs = getattr(model, "models", None)
    if subs is None:
        raise TypeError(f"expected a model list with a '.models' attribute; got {type(model)}")
    out = []
    for gp in subs:
        out.append(gp if isinstance(gp, GPObjective) else _extract_single(gp))
    return GPModelList(out)


def _softplus(x: float) -> float:
    if x == -math.inf:
        return 0.0
    return math.log1p(math.exp(-abs(x))) + max(x, 0.0)


def model_from_problem_blob(
    blob: dict, noise: Optional[Sequence[float]] = None, kernel: int = MATERN52
) -> GPModelList:
    """Surrogate the reference builds with ``--fit-hyperparams=never`` (``bo_loop.py:574-589``)
    from a committed problem dict (``data_catalog.py:99-111``: ``train_x``, ``train_y``,
    ``fixed_hyperparams``, ``model_state_dict``).  ``noise`` defaults to the likelihood noise in
    the state dict (``softplus(raw) + lower_bound``; 1e-8 in the committed files)."""
    hp = blob["fixed_hyperparams"]
    sd = blob.get("model_state_dict", {})
    tx = torch.as_tensor(blob["train_x"], dtype=torch.double)
    ty = torch.as_tensor(blob["train_y"], dtype=torch.double)
    objs = []
    for m in range(ty.shape[1]):
        if noise is not None:
            nz = float(noise[m])
        else:
            raw = float(sd[f"models.{m}.likelihood.noise_covar.raw_noise"].reshape(-1)[0])
            lb = float(sd[f"models.{m}.likelihood.noise_covar.raw_noise_constraint.lower_bound"])
            nz = _softplus(raw) + lb
        objs.append(
            GPObjective(
                train_x=tx.clone(),
                train_y=ty[:, m].clone(),
                lengthscale=torch.full((tx.shape[1],), float(hp["length_scales"][m])),
                outputscale=float(hp["output_scales"][m]),
                mean_const=float(hp["means"][m]),
                noise=nz,
                kernel=kernel,
            )
        )
    return GPModelList(objs)
