"""Small discretisations take the single-CTA-per-candidate path (small_kg_kernel, DKG_SMALL=0 disables
it): it must agree with the staged pipeline (int8 contraction, filters, warp hull) and the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import oracle_model, small_problem
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _run(P, target, X, small):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    os.environ["DKG_SMALL"] = "1" if small else "0"
    try:
        acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=target)
        Xg = X.clone().requires_grad_(True)
        kg = acq(Xg.unsqueeze(1))
        (g,) = torch.autograd.grad(kg.sum(), Xg)
        st = acq._get_plan().stats()
        return kg.detach(), g, st
    finally:
        os.environ.pop("DKG_SMALL", None)


@pytest.mark.parametrize("kw", [
    dict(),
    dict(kernel=1, d=3, seed=1),
    dict(y_std=(2.5, 0.5), y_mean=(1.0, -3.0), seed=2),
    dict(n_train_per_obj=(24, 17), n_train=24, seed=3),
    dict(d=2, n_train=60, n_disc=121, n_scal=16, n_cand=10, seed=4),   # the reference's BO-loop preset shape
    dict(d=1, n_train=8, n_disc=255, n_scal=3, n_cand=5, seed=5, lengthscales=(0.4, 0.6)),  # largest N, long hulls
    dict(d=8, n_train=40, n_disc=200, n_scal=20, n_cand=7, seed=6),
], ids=["matern-d2", "rbf-d3", "standardised", "ragged", "bo-preset", "N255-d1", "d8-S20"])
def test_small_path_matches_staged_pipeline_and_oracle(kw):
    P = small_problem(**kw)
    om = oracle_model(P.model)
    for target in (0, 1):
        kg_s, g_s, st_s = _run(P, target, P.candidates, small=True)
        kg_b, g_b, st_b = _run(P, target, P.candidates, small=False)
        scale = float(torch.cat([o.train_y for o in P.model.models]).abs().max())
        np.testing.assert_allclose(kg_s.numpy(), kg_b.numpy(), rtol=1e-9, atol=1e-12 * scale)
        np.testing.assert_allclose(g_s.numpy(), g_b.numpy(), rtol=1e-6, atol=1e-10 * scale)
        assert st_s[3] == st_b[3]  # same hulls
        Xo = P.candidates.clone().requires_grad_(True)
        want = odk.forward(om, Xo.unsqueeze(1), P.x_disc, P.weights, target, dense=False)
        want.sum().backward()
        np.testing.assert_allclose(kg_s.numpy(), want.detach().numpy(), rtol=1e-9, atol=1e-12 * scale)
        np.testing.assert_allclose(g_s.numpy(), Xo.grad.numpy(), rtol=1e-6, atol=1e-10 * scale)
        assert int(kg_s.argmax()) == int(want.argmax())


def test_small_path_is_batch_composition_independent_and_deterministic():
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    P = small_problem(n_disc=121, n_cand=40)
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=1)
    with torch.no_grad():
        full = acq(P.candidates.unsqueeze(1))
        again = acq(P.candidates.unsqueeze(1))
        sub = acq(P.candidates[7:19].unsqueeze(1))
        one = acq(P.candidates[11:12].unsqueeze(1))
    assert torch.equal(full, again) and torch.equal(sub, full[7:19]) and torch.equal(one, full[11:12])
    # device tensors and host tensors give the same bits (graph replay over the staging buffers either way)
    with torch.no_grad():
        dev = DiscreteKnowledgeGradient(P.model, P.x_disc.cuda(), P.weights, target_output_ix=1)(P.candidates.cuda().unsqueeze(1))
    assert torch.equal(dev.cpu(), full)
