"""The handful of BoTorch names the hot path's public surface uses (``discretekg.py:12-20``).

If BoTorch is importable (the production pipeline), the real classes are re-exported so that
``DiscreteKnowledgeGradient`` IS a ``botorch.acquisition.AcquisitionFunction`` and raises the real
``BotorchTensorDimensionError`` / ``UnsupportedError``.  Otherwise (this build container, the GPU
test boxes) minimal local equivalents with the same names and behaviour are used.
"""

from __future__ import annotations

import functools

import torch
from torch import Tensor

try:  # pragma: no cover - exercised only where botorch is installed
    from botorch.acquisition import AcquisitionFunction
    from botorch.exceptions import BotorchTensorDimensionError, UnsupportedError
    from botorch.utils import draw_sobol_samples, t_batch_mode_transform

    HAVE_BOTORCH = True
except ImportError:
    HAVE_BOTORCH = False

    class BotorchError(Exception):
        pass

    class BotorchTensorDimensionError(BotorchError):
        pass

    class UnsupportedError(BotorchError):
        pass

    class AcquisitionFunction(torch.nn.Module):
        """Abstract base: stores ``model`` (``botorch/acquisition/acquisition.py``)."""

        def __init__(self, model) -> None:
            super().__init__()
            # GP state containers are plain objects, not nn.Modules
            object.__setattr__(self, "model", model)

        def set_X_pending(self, X_pending=None) -> None:
            raise NotImplementedError

    def draw_sobol_samples(bounds: Tensor, n: int, q: int, seed=None) -> Tensor:
        """``n x q x d`` scrambled-Sobol points inside ``bounds`` (``2 x d``)."""
        d = bounds.shape[-1]
        engine = torch.quasirandom.SobolEngine(q * d, scramble=True, seed=seed)
        u = engine.draw(n, dtype=bounds.dtype).view(n, q, d).to(bounds.device)
        return bounds[0] + (bounds[1] - bounds[0]) * u

    def t_batch_mode_transform(expected_q=None):
        """Adds a t-batch dimension to 2-D inputs and checks ``q`` (and the output shape)."""

        def decorator(method):
            @functools.wraps(method)
            def decorated(acqf, X, *args, **kwargs):
                if X.dim() < 2:
                    raise ValueError(
                        f"{type(acqf).__name__} requires X to have at least 2 dimensions,"
                        f" but received X with only {X.dim()} dimensions."
                    )
                if expected_q is not None and X.shape[-2] != expected_q:
                    raise AssertionError(
                        f"Expected X to be `batch_shape x q={expected_q} x d`, but"
                        f" got X with shape {X.shape}."
                    )
                X = X if X.dim() > 2 else X.unsqueeze(0)
                return method(acqf, X, *args, **kwargs)

            return decorated

        return decorator
