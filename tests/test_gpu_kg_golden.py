"""DiscreteKnowledgeGradient (CUDA, through the C-ABI) against the golden KG values produced by
the reference's own forward code on the oracle GP, and stage-by-stage against the oracle."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_KG_SPECS, load_golden, oracle_model, small_problem
from oracle import discretekg as odk
from oracle import gp as ogp

pytestmark = pytest.mark.gpu


def _acqf(P, target, device=None):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    xd = P.x_disc if device is None else P.x_disc.to(device)
    return DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=target)


@pytest.mark.parametrize("name", sorted(GOLDEN_KG_SPECS))
@pytest.mark.parametrize("target", [0, 1, None])
def test_kg_and_gradient_match_reference_code_golden(name, target):
    G = load_golden("kg_reference_code_golden.npz")
    P = small_problem(**GOLDEN_KG_SPECS[name])
    acq = _acqf(P, target)
    X = P.candidates.clone().requires_grad_(True)
    kg = acq(X.unsqueeze(1))
    (g,) = torch.autograd.grad(kg.sum(), X)
    key = "coupled" if target is None else f"t{target}"
    want, want_g = G[f"{name}__{key}__kg"], G[f"{name}__{key}__grad"]
    scale = float(acq._get_plan().read("A0").abs().max())
    # fp64 mode: rel 1e-9 (north star), with an absolute floor of 1e-12 x |intercepts| because KG
    # is the small difference E[max] - max
    np.testing.assert_allclose(kg.detach().numpy(), want, rtol=1e-9, atol=1e-12 * scale)
    np.testing.assert_allclose(g.numpy(), want_g, rtol=1e-7, atol=1e-10 * scale)
    assert int(np.argmax(kg.detach().numpy())) == int(np.argmax(want))


def test_stages_match_oracle():
    P = small_problem(d=3, n_train=40, n_disc=130, n_scal=5, n_cand=20, seed=7)
    om = oracle_model(P.model, distance="direct")
    for target in (0, 1):
        acq = _acqf(P, target)
        with torch.no_grad():
            acq(P.candidates.unsqueeze(1))
        plan = acq._get_plan()
        o = om.models[target]
        cache = ogp._train_cache(o)
        L = plan.read("chol").cpu()
        np.testing.assert_allclose(L.numpy(), cache["L"].numpy(), rtol=1e-10, atol=1e-12)
        Kxd = ogp.kernel_matrix(o, o.train_x, P.x_disc)
        B = torch.cholesky_solve(Kxd, cache["L"])
        np.testing.assert_allclose(plan.read("B").cpu().numpy(), B.numpy(), rtol=1e-7, atol=1e-9)
        mu = torch.stack([ogp.posterior(m, P.x_disc)[0] for m in om.models], dim=-1)
        np.testing.assert_allclose(plan.read("mu_disc").cpu().numpy(), mu.numpy(), rtol=1e-9, atol=1e-11)
        for c in (0, 7, 19):
            a, b = odk.lines_single_output(om, P.candidates[c], target, P.x_disc, P.weights, dense=True)
            sl = plan.read("slopes")[c].cpu()
            j = int(torch.argmax(P.weights[:, target]))
            z = b[j] / P.weights[j, target]
            np.testing.assert_allclose(sl[:-1].numpy(), z[1:].numpy(), rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(float(sl[-1]), float(z[0]), rtol=1e-8, atol=1e-11)
            np.testing.assert_allclose(plan.read("a_new")[c].cpu().numpy(), a[:, 0].numpy(), rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(plan.read("A0").cpu().numpy(), a[:, 1:].numpy(), rtol=1e-10, atol=1e-12)


def test_device_and_host_inputs_agree_and_tbatch_shapes():
    P = small_problem(n_cand=24)
    acq = _acqf(P, 0)
    X = P.candidates.reshape(2, 12, 1, P.d)
    with torch.no_grad():
        kg_host = acq(X)
        kg_dev = acq(X.cuda())
        kg_2d = acq(P.candidates[:1])  # (1, d) -> t-batch of one
    assert kg_host.shape == (2, 12) and kg_dev.shape == (2, 12) and kg_dev.is_cuda
    assert torch.equal(kg_host, kg_dev.cpu())
    assert kg_2d.shape == (1,) and kg_2d[0] == kg_host[0, 0]
    empty = acq(torch.zeros(0, 1, P.d))
    assert empty.shape == (0,)


def test_gradcheck_like_reference():  # test_discretekg.py:110-135, at x = (0.51, 0.51)
    P = small_problem(n_disc=9, n_train=10, n_scal=3)
    P.x_disc = torch.stack(torch.meshgrid(torch.linspace(0, 1, 3), torch.linspace(0, 1, 3), indexing="ij"), -1).reshape(-1, 2).double()
    for target in (0, 1):
        acq = _acqf(P, target)
        x = torch.tensor([[[0.51, 0.51]]], dtype=torch.double, requires_grad=True)
        torch.autograd.gradcheck(acq, (x,), raise_exception=True, eps=1e-6, atol=1e-6, rtol=1e-4)


def test_negative_and_zero_weights():
    P = small_problem(n_scal=4)
    P.weights = torch.tensor([[0.0, 1.0], [1.0, 0.0], [-0.5, 1.5], [0.3, 0.7]], dtype=torch.double)
    om = oracle_model(P.model)
    for target in (0, 1):
        acq = _acqf(P, target)
        X = P.candidates.clone().requires_grad_(True)
        kg = acq(X.unsqueeze(1))
        (g,) = torch.autograd.grad(kg.sum(), X)
        Xo = P.candidates.clone().requires_grad_(True)
        want = odk.forward(om, Xo.unsqueeze(1), P.x_disc, P.weights, target, dense=True)
        (go,) = torch.autograd.grad(want.sum(), Xo)
        np.testing.assert_allclose(kg.detach().numpy(), want.detach().numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(g.numpy(), go.numpy(), rtol=1e-7, atol=1e-11)


def test_candidate_on_training_point_and_outside_unit_cube():
    P = small_problem(noise=(1e-4, 1e-4))
    om = oracle_model(P.model)
    X = torch.cat([P.model.models[0].train_x[:3], torch.tensor([[1.3, -0.2], [0.0, 1.0]], dtype=torch.double)])
    acq = _acqf(P, 0)
    with torch.no_grad():
        kg = acq(X.unsqueeze(1))
    want = odk.forward(om, X.unsqueeze(1), P.x_disc, P.weights, 0, dense=True)
    np.testing.assert_allclose(kg.numpy(), want.numpy(), rtol=1e-7, atol=1e-11)


def test_posterior_mean_entry_point():
    """dkg_posterior_mean_dev == model.posterior(X).mean (oracle), incl. Standardize and ragged n."""
    for name in ("std_d2", "ragged_d2", "rbf_d3"):
        P = small_problem(**GOLDEN_KG_SPECS[name])
        om = oracle_model(P.model)
        plan = _acqf(P, 0)._get_plan()
        X = torch.rand(37, P.d, dtype=torch.double)
        got = plan.posterior_mean(X)
        want = torch.stack([ogp.posterior(o, X)[0] for o in om.models], dim=-1)
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-10, atol=1e-12)
