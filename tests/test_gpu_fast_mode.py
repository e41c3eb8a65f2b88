"""Reduced-precision mode (DKG_PLAN_FAST32 / ``acq.precision = "float32"``): the covariance
contraction keeps 4 base-256 digits per operand.  Stated tolerance against the float64 oracle:
|dKG| <= 1e-4 |KG| + 1e-7 max|intercept|; the argmax candidate must still agree when the two best
values are separated by more than that tolerance."""
import numpy as np
import pytest
import torch

from helpers import oracle_model
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _check(P, n_oracle):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    om = oracle_model(P.model)
    for target in (0, 1):
        acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=target)
        with torch.no_grad():
            kg64 = acq(P.candidates.unsqueeze(1))
        acq.precision = "float32"
        assert acq._plan is None  # the cached GPU state is rebuilt for the new mode
        X = P.candidates.clone().requires_grad_(True)
        kg32 = acq(X.unsqueeze(1))
        (g32,) = torch.autograd.grad(kg32.sum(), X)
        assert acq._get_plan().stats()[7] == 1  # int8 tensor-core contraction
        scale = float(acq._get_plan().read("A0").abs().max())
        tol = 1e-4 * kg64.abs() + 1e-7 * scale
        assert bool(((kg32.detach() - kg64).abs() <= tol).all())
        assert not torch.equal(kg32.detach(), kg64)  # the mode really computes something else
        assert torch.isfinite(g32).all()
        want = odk.forward(om, P.candidates[:n_oracle].unsqueeze(1), P.x_disc, P.weights, target, dense=False)
        np.testing.assert_allclose(kg32.detach()[:n_oracle].numpy(), want.numpy(), rtol=1e-4, atol=1e-7 * scale)
        top2 = torch.topk(kg64, 2).values
        if float(top2[0] - top2[1]) > 2 * float(tol.max()):
            assert int(kg32.argmax()) == int(kg64.argmax())


def test_fast_mode_c2_shape():
    from decoupledbo_b200 import synthetic

    P = synthetic.problem_c2(n_cand=48)
    for o, nz in zip(P.model.models, (1e-2, 0.5)):  # the well-conditioned variant (see test_gpu_full_size)
        o.noise = nz
    _check(P, 48)


def test_fast_mode_c4_shape():
    from decoupledbo_b200 import synthetic

    _check(synthetic.problem_c4(n_cand=256), 4)


def test_precision_validation():
    from decoupledbo_b200 import synthetic
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    P = synthetic.problem_c2(n_cand=4)
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)
    assert acq.precision == "float64"
    with pytest.raises(ValueError):
        acq.precision = "bfloat16"
