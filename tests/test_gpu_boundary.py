"""The module-level surface of the reference's discretekg.py on the drop-in (GPU-backed), beyond
the reference's own KATs (tests/golden/reference_suite): golden line sets, random sets against the
oracle, hulls of more than 64 vertices through the KG backward, constructor helpers and the
model-change fingerprint."""
import os

import numpy as np
import pytest
import torch

from helpers import load_golden, oracle_model, small_problem
from oracle import discretekg as odk


def _dk():
    from decoupledbo_b200.modules.acquisition import discretekg

    return discretekg


# ---- host-side validation: same exception types as the reference (no GPU needed) --------------
def test_verify_raises_botorch_dimension_errors():
    dk = _dk()
    from decoupledbo_b200.botorch_compat import BotorchTensorDimensionError

    with pytest.raises(BotorchTensorDimensionError, match="one-dimensional"):  # discretekg.py:456-460
        dk.calculate_epigraph_indices(torch.zeros(2, 2), torch.zeros(2, 2))
    with pytest.raises(BotorchTensorDimensionError, match="same shape"):  # :461-465
        dk.calculate_epigraph_indices(torch.zeros(3), torch.zeros(2))
    with pytest.raises(ValueError, match="at least one line"):  # :466-470
        dk._verify_intercepts_and_slopes(torch.zeros(0), torch.zeros(0))
    with pytest.raises(BotorchTensorDimensionError, match="boundaries"):  # :425-429
        dk.calculate_expected_value_of_piecewise_linear_function(torch.zeros(3), torch.zeros(3), torch.zeros(3))
    with pytest.raises(BotorchTensorDimensionError, match="two dimensions"):  # :175-180, :259-264
        dk.calculate_discrete_kg(None, torch.zeros(2), torch.zeros(4, 2), torch.zeros(2))
    with pytest.raises(BotorchTensorDimensionError, match="two dimensions"):
        dk.calculate_discrete_kg_conditioning_on_single_output(None, torch.zeros(2), 0, torch.zeros(4, 2), torch.zeros(2))
    from decoupledbo_b200.botorch_compat import UnsupportedError

    with pytest.raises(UnsupportedError, match="ModelListGP"):  # :270-273
        dk.calculate_discrete_kg_conditioning_on_single_output(object(), torch.zeros(2), 0, torch.zeros(4, 2), torch.zeros(1, 2))


def test_create_with_sobol_sample_builds_the_discretisation():  # discretekg.py:33-60
    dk = _dk()
    P = small_problem()
    bounds = torch.tensor([[0.0, -1.0], [2.0, 1.0]], dtype=torch.double)
    acq = dk.DiscreteKnowledgeGradient.create_with_sobol_sample(
        P.model, bounds, num_discrete_points=37, scalarisation_weights=P.weights, target_output_ix=1)
    xd = acq.x_discretisation
    assert xd.shape == (37, 2) and xd.dtype == bounds.dtype and xd.device == bounds.device
    assert bool((xd >= bounds[0]).all()) and bool((xd <= bounds[1]).all())
    assert acq.target_output_ix == 1 and acq.scalarisation_weights is P.weights and acq.model is P.model
    assert len(torch.unique(xd, dim=0)) == 37  # a Sobol sample, not a repeated point


@pytest.mark.gpu
def test_create_with_sobol_sample_forward_matches_explicit_discretisation():
    dk = _dk()
    P = small_problem()
    bounds = torch.tensor([[0.0, 0.0], [1.0, 1.0]], dtype=torch.double)
    acq = dk.DiscreteKnowledgeGradient.create_with_sobol_sample(P.model, bounds, 50, P.weights, target_output_ix=0)
    ref = dk.DiscreteKnowledgeGradient(P.model, acq.x_discretisation.clone(), P.weights, target_output_ix=0)
    with torch.no_grad():
        a, b = acq(P.candidates.unsqueeze(1)), ref(P.candidates.unsqueeze(1))
    assert torch.equal(a, b)
    want = odk.forward(oracle_model(P.model), P.candidates.unsqueeze(1), acq.x_discretisation, P.weights, 0, dense=True)
    np.testing.assert_allclose(a.numpy(), want.numpy(), rtol=1e-9, atol=1e-12)


# ---- calculate_epigraph_indices / expectation on the GPU ---------------------------------------
@pytest.mark.gpu
def test_epigraph_indices_golden_sets_bit_exact():
    dk = _dk()
    G = load_golden("epigraph_golden.npz")
    for k in range(int(G["n_sets"])):
        a, b = torch.tensor(G[f"a{k}"]), torch.tensor(G[f"b{k}"])
        idx, x = dk.calculate_epigraph_indices(a, b)
        want_idx = G[f"idx{k}"].tolist()
        assert idx.dtype == torch.long and x.dtype == torch.double
        np.testing.assert_array_equal(x.numpy(), G[f"x{k}"], err_msg=str(k))
        if idx.tolist() != want_idx:  # only identical duplicate lines may be swapped
            for i, j in zip(idx.tolist(), want_idx):
                assert a[i] == a[j] and b[i] == b[j]
        e = dk.calculate_expected_value_of_piecewise_linear_function(a[idx], b[idx], x)
        assert abs(float(e) - float(G[f"e{k}"])) <= 1e-15 * max(1.0, float(a.abs().max()))


@pytest.mark.gpu
def test_epigraph_with_more_than_64_vertices_and_autograd_of_intersections():
    dk = _dk()
    L = 500
    b = torch.linspace(-3, 3, L, dtype=torch.double).requires_grad_(True)
    a = (-0.5 * b.detach() ** 2).requires_grad_(True)
    idx, x = dk.calculate_epigraph_indices(a, b)
    idx_o, x_o = odk.epigraph_indices(a.detach(), b.detach())
    assert idx.tolist() == idx_o.tolist() and len(idx) == L
    np.testing.assert_array_equal(x.detach().numpy(), x_o.numpy())
    (ga,) = torch.autograd.grad(x.sum(), a)  # intersections carry the reference's autograd graph
    assert ga.shape == (L,) and torch.isfinite(ga).all()


@pytest.mark.gpu
def test_piecewise_expectation_arbitrary_boundaries_and_gradients():
    dk = _dk()
    rng = np.random.default_rng(5)
    for H in (1, 2, 5, 33, 70):
        a = torch.tensor(rng.normal(size=H), requires_grad=True)
        b = torch.tensor(rng.normal(size=H), requires_grad=True)
        z = torch.tensor(np.sort(rng.normal(size=H - 1)), requires_grad=True)  # NOT the lines' intersections
        e = dk.calculate_expected_value_of_piecewise_linear_function(a, b, z)
        ao, bo, zo = (t.detach().clone().requires_grad_(True) for t in (a, b, z))
        eo = odk.expected_value_of_piecewise_linear_function(ao, bo, zo)
        assert abs(float(e) - float(eo)) <= 4e-16 * max(1.0, float(a.abs().sum() + b.abs().sum()))
        e.backward()
        eo.backward()
        np.testing.assert_allclose(a.grad.numpy(), ao.grad.numpy(), rtol=0, atol=1e-15)
        np.testing.assert_allclose(b.grad.numpy(), bo.grad.numpy(), rtol=0, atol=1e-15)
        if H > 1:
            np.testing.assert_allclose(z.grad.numpy(), zo.grad.numpy(), rtol=1e-13, atol=1e-15)


# ---- more than 64 hull vertices through the KG backward ----------------------------
def _smooth_1d_problem(n_disc=4096):
    from decoupledbo_b200.gp_state import GPModelList, GPObjective

    g = torch.Generator().manual_seed(3)
    tx = torch.rand(6, 1, generator=g, dtype=torch.double)
    objs = [GPObjective(train_x=tx, train_y=torch.randn(6, generator=g, dtype=torch.double),
                        lengthscale=torch.tensor([0.35 + 0.2 * m]), outputscale=1.0 + m, mean_const=0.0,
                        noise=1e-2, kernel=1) for m in range(2)]
    disc = torch.linspace(0, 1, n_disc, dtype=torch.double).unsqueeze(1)
    W = torch.tensor([[0.7, 0.3], [0.2, 0.8]], dtype=torch.double)
    X = torch.tensor([[0.31], [0.62], [0.05]], dtype=torch.double)
    return GPModelList(objs), disc, W, X


@pytest.mark.gpu
@pytest.mark.parametrize("target", [0, None], ids=["decoupled", "coupled"])
def test_gradient_with_hundreds_of_hull_vertices(target):
    dk = _dk()
    model, disc, W, X = _smooth_1d_problem()
    om = oracle_model(model)
    acq = dk.DiscreteKnowledgeGradient(model, disc, W, target_output_ix=target)
    Xg = X.clone().requires_grad_(True)
    kg = acq(Xg.unsqueeze(1))
    kg.sum().backward()
    stats = acq._get_plan().stats()
    assert stats[3] > 64 * X.shape[0] * W.shape[0], stats  # on average > 64 vertices per set
    assert stats[6] == 0
    Xo = X.clone().requires_grad_(True)
    want = odk.forward(om, Xo.unsqueeze(1), disc, W, target, dense=False)
    want.sum().backward()
    np.testing.assert_allclose(kg.detach().numpy(), want.detach().numpy(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(Xg.grad.numpy(), Xo.grad.numpy(), rtol=1e-6, atol=1e-11)
    # the same through CUDA tensors (no host-side check on that path)
    Xd = X.cuda().requires_grad_(True)
    acq_d = dk.DiscreteKnowledgeGradient(model, disc.cuda(), W, target_output_ix=target)
    acq_d(Xd.unsqueeze(1)).sum().backward()
    np.testing.assert_allclose(Xd.grad.cpu().numpy(), Xo.grad.numpy(), rtol=1e-6, atol=1e-11)


@pytest.mark.gpu
def test_exhausted_spill_pool_fails_loudly(monkeypatch):
    dk = _dk()
    model, disc, W, X = _smooth_1d_problem()
    monkeypatch.setenv("DKG_SPILL_BLOCKS", "1")  # room for 32 extra vertices in total
    acq = dk.DiscreteKnowledgeGradient(model, disc, W, target_output_ix=0)
    with pytest.raises(RuntimeError, match="DKG_SPILL_BLOCKS"):  # host tensors: checked for free
        acq(X.clone().requires_grad_(True).unsqueeze(1)).sum().backward()
    acq_d = dk.DiscreteKnowledgeGradient(model, disc.cuda(), W, target_output_ix=0)
    Xd = X.cuda().requires_grad_(True)
    kg = acq_d(Xd.unsqueeze(1))  # device tensors: no sync, the gradient rows carry NaN instead
    kg.sum().backward()
    assert torch.isfinite(kg).all()  # the values stay exact
    assert torch.isnan(Xd.grad).any()
    assert acq_d._get_plan().stats()[6] > 0


# ---- target index, stale-cache protection, grad gating ----------------------------------------
@pytest.mark.gpu
def test_negative_target_index_counts_from_the_end():
    dk = _dk()
    P = small_problem()
    X = P.candidates.unsqueeze(1)
    with torch.no_grad():
        last = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=1)(X)
        neg = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=-1)(X)
        coupled = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights)(X)
    assert torch.equal(last, neg) and not torch.equal(neg, coupled)
    with pytest.raises(IndexError):
        dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=2)(X)


@pytest.mark.gpu
def test_plan_follows_model_changes():
    dk = _dk()
    P = small_problem()
    X = P.candidates.unsqueeze(1)
    acq = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)
    with torch.no_grad():
        before = acq(X)
        plan0 = acq._plan
        assert acq(X) is not None and acq._plan is plan0  # unchanged model: the plan is reused
        P.model.models[0].noise *= 3.0  # hyper-parameter change
        after = acq(X)
        fresh = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)(X)
        assert acq._plan is not plan0 and torch.equal(after, fresh) and not torch.equal(after, before)
        P.model.models[1].train_y.mul_(1.5)  # in-place data change
        assert torch.equal(acq(X), dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)(X))


@pytest.mark.gpu
def test_module_level_kg_functions_match_the_class_and_the_oracle():
    dk = _dk()
    P = small_problem()
    om = oracle_model(P.model)
    for c in range(3):
        x = P.candidates[c].clone().requires_grad_(True)
        v = dk.calculate_discrete_kg_conditioning_on_single_output(P.model, x, 1, P.x_disc, P.weights)
        v.backward()
        xo = P.candidates[c].clone().requires_grad_(True)
        vo = odk.kg_single_output(om, xo, 1, P.x_disc, P.weights)
        vo.backward()
        assert v.dim() == 0 and float(v) == pytest.approx(float(vo), rel=1e-9, abs=1e-13)
        np.testing.assert_allclose(x.grad.numpy(), xo.grad.numpy(), rtol=1e-6, atol=1e-11)
        vc = dk.calculate_discrete_kg(P.model, P.candidates[c], P.x_disc, P.weights)
        assert float(vc) == pytest.approx(float(odk.kg_coupled(om, P.candidates[c], P.x_disc, P.weights)), rel=1e-9, abs=1e-13)


@pytest.mark.gpu
def test_no_grad_evaluation_of_a_leaf_skips_the_backward():
    dk = _dk()
    from decoupledbo_b200 import _native

    P = small_problem()
    acq = dk.DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)
    X = P.candidates.clone().requires_grad_(True).unsqueeze(1)
    with torch.no_grad():
        kg = acq(X)
    assert not kg.requires_grad
    assert acq._get_plan().read("kg_terms").shape == (P.candidates.shape[0], P.weights.shape[0])
