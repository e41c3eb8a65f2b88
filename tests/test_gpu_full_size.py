"""BASELINE.json full sizes (c2 and c4 shapes): size-independent properties, plus oracle spot
checks on a few candidates (the row-only oracle costs ~0.5 s per candidate at c4)."""
import numpy as np
import pytest
import torch

from helpers import oracle_model
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _acqf(P, target, xd=None):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    return DiscreteKnowledgeGradient(P.model, P.x_disc if xd is None else xd, P.weights, target_output_ix=target)


@pytest.fixture(scope="module")
def c4():
    from decoupledbo_b200 import synthetic

    return synthetic.problem_c4(n_cand=1024)


def test_c4_properties_and_spot_parity(c4):
    P = c4
    dev = torch.device("cuda")
    X = P.candidates.to(dev)
    om = oracle_model(P.model)
    for target in (0, 1):
        acq = _acqf(P, target, P.x_disc.to(dev))
        Xg = X.clone().requires_grad_(True)
        kg = acq(Xg.unsqueeze(1))
        (g,) = torch.autograd.grad(kg.sum(), Xg)
        assert torch.isfinite(kg).all() and torch.isfinite(g).all()
        scale = float(acq._get_plan().read("A0").abs().max())
        assert float(kg.min()) >= -1e-12 * scale  # KG is non-negative up to rounding
        # idempotence / determinism: identical bits on a second evaluation
        with torch.no_grad():
            kg2 = acq(X.unsqueeze(1))
        assert torch.equal(kg.detach(), kg2)
        # batch-composition independence: a sub-batch gives the same bits as the full batch
        with torch.no_grad():
            sub = acq(X[100:163].unsqueeze(1))
        assert torch.equal(sub, kg2[100:163])
        # permuting the discretisation leaves KG unchanged (up to summation order in the means)
        perm = torch.randperm(P.x_disc.shape[0], generator=torch.Generator().manual_seed(1))
        acq_p = _acqf(P, target, P.x_disc[perm].to(dev))
        with torch.no_grad():
            kg_p = acq_p(X[:256].unsqueeze(1))
        np.testing.assert_allclose(kg_p.cpu().numpy(), kg2[:256].cpu().numpy(), rtol=1e-8, atol=1e-12 * scale)
        # oracle spot checks (row-only posterior), values + gradients + argmax over the spots
        spots = [0, 1, 2, 3, 599, 977, 1023]
        want, want_g = [], []
        for c in spots:
            x = P.candidates[c].clone().requires_grad_(True)
            v = odk.kg_single_output(om, x, target, P.x_disc, P.weights, dense=False)
            v.backward()
            want.append(v.item())
            want_g.append(x.grad.numpy())
        got = kg.detach().cpu().numpy()[spots]
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12 * scale)
        np.testing.assert_allclose(g.cpu().numpy()[spots], np.array(want_g), rtol=1e-6, atol=1e-10 * scale)
        assert int(np.argmax(got)) == int(np.argmax(want))


def test_c2_full_parity_and_objective_choice():
    from decoupledbo_b200 import synthetic
    from decoupledbo_b200.modules.acquisition_optimisation_strategy import choose_best_objective

    P = synthetic.problem_c2(n_cand=64)
    om = oracle_model(P.model)
    best, best_o = [], []
    for target in (0, 1):
        acq = _acqf(P, target)
        with torch.no_grad():
            kg = acq(P.candidates.unsqueeze(1))
        want = odk.forward(om, P.candidates.unsqueeze(1), P.x_disc, P.weights, target, dense=False)
        scale = float(acq._get_plan().read("A0").abs().max())
        np.testing.assert_allclose(kg.numpy(), want.numpy(), rtol=1e-9, atol=1e-12 * scale)
        assert int(kg.argmax()) == int(want.argmax())  # bit-exact argmax candidate
        i = int(kg.argmax())
        best.append((target, P.candidates[i : i + 1], kg[i]))
        best_o.append((target, P.candidates[i : i + 1], want[i]))
    costs = [1.0, 1.0]
    assert choose_best_objective(best, costs)[0] == odk.choose_best_objective(best_o, costs)[0]
