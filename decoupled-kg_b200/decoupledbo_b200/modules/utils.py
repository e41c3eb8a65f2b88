"""Host-side helpers of the hot path's callers (mirror of
``src/decoupledbo/modules/utils.py`` for the one function the path uses)."""

import torch


def make_torch_std_grid(n_points_per_axis, n_dimensions, tkwargs=None):
    """``n_points_per_axis**d x d`` points of a regular grid on [0, 1]^d.

    Same contract as the reference's ``make_torch_std_grid`` (``utils.py:79-107``): row-major
    enumeration with the FIRST coordinate varying slowest, e.g. for (3, 2):
    (0,0) (0,.5) (0,1) (.5,0) ... (1,1).  ``tkwargs`` are forwarded to ``torch.linspace``.
    """
    tkwargs = tkwargs or {}
    if n_dimensions <= 0:
        raise ValueError(f"Expected n_dimensions >= 1. Got {n_dimensions}.")
    axis = torch.linspace(0, 1, n_points_per_axis, **tkwargs)
    mesh = torch.meshgrid(*([axis] * n_dimensions), indexing="ij")
    return torch.stack([m.reshape(-1) for m in mesh], dim=-1)

