"""GPU parity (through the C-ABI) against the oracle at sizes the oracle finishes in seconds."""
import numpy as np
import pytest
import torch

from helpers import oracle_model, small_problem
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _acqf(P, target):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    return DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=target)


@pytest.mark.parametrize("target", [0, 1])
@pytest.mark.parametrize("kernel", [0, 1])
def test_forward_matches_oracle(target, kernel):
    P = small_problem(kernel=kernel)
    om = oracle_model(P.model)
    acq = _acqf(P, target)
    with torch.no_grad():
        kg = acq(P.candidates.unsqueeze(1))
    want = odk.forward(om, P.candidates.unsqueeze(1), P.x_disc, P.weights, target, dense=True)
    scale = float(torch.max(torch.abs(acq._get_plan().read("A0"))))
    # fp64 tolerance: rel 1e-9 on KG, plus an absolute floor for the E[max]-max cancellation
    np.testing.assert_allclose(kg.numpy(), want.numpy(), rtol=1e-9, atol=1e-12 * scale)
    assert int(torch.argmax(kg)) == int(torch.argmax(want))
