import functools

import torch


def draw_sobol_samples(bounds, n, q, seed=None):
    d = bounds.shape[-1]
    eng = torch.quasirandom.SobolEngine(q * d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=bounds.dtype).view(n, q, d)
    return bounds[0] + (bounds[1] - bounds[0]) * u


def t_batch_mode_transform(expected_q=None):
    def deco(method):
        @functools.wraps(method)
        def wrapped(self, X, *args, **kwargs):
            if X.dim() < 2:
                raise ValueError("X must have at least 2 dimensions")
            if expected_q is not None and X.shape[-2] != expected_q:
                raise AssertionError(f"Expected X to be `batch_shape x q={expected_q} x d`")
            if X.dim() == 2:
                X = X.unsqueeze(0)
            return method(self, X, *args, **kwargs)

        return wrapped

    return deco
