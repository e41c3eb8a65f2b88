"""The vectorised CPU baseline (oracle/vectorised.py, BASELINE.md section 3 "(V)") against the
reference-faithful oracle: same KG values and gradients to rounding."""
import numpy as np
import pytest
import torch

from helpers import oracle_model, small_problem
from oracle import discretekg as odk
from oracle import vectorised as ov


@pytest.mark.parametrize("kw", [dict(), dict(kernel=1, d=3, seed=1), dict(y_std=(2.5, 0.5), y_mean=(1.0, -3.0), seed=2),
                                dict(n_train_per_obj=(24, 17), n_train=24, seed=3)])
def test_vectorised_baseline_matches_the_oracle(kw):
    P = small_problem(**kw)
    om = oracle_model(P.model)
    for target in (0, 1):
        prep = ov.Prepared(om, P.x_disc, P.weights, target)
        kg, dX = ov.kg_batch(prep, P.candidates, need_grad=True, threads=2)
        X = P.candidates.clone().requires_grad_(True)
        want = odk.forward(om, X.unsqueeze(1), P.x_disc, P.weights, target, dense=True)
        want.sum().backward()
        np.testing.assert_allclose(kg.numpy(), want.detach().numpy(), rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(dX.numpy(), X.grad.numpy(), rtol=1e-7, atol=1e-11)
        kg2, none = ov.kg_batch(prep, P.candidates, need_grad=False, threads=1)
        assert none is None and torch.equal(kg2, kg)
