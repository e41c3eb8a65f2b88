"""Incremental plan refresh (dkg_plan_append_point): 50 sequential single-point appends,
ragged across objectives, must track a freshly built plan -- values within the path's stated tolerance
(rel 1e-9 + 1e-12 max|intercept|; the observed maximum is far smaller), gradients rel 1e-6."""
import numpy as np
import pytest
import torch

from helpers import small_problem

pytestmark = pytest.mark.gpu


def _fresh(P, target, xd=None):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    return DiscreteKnowledgeGradient(P.model, P.x_disc if xd is None else xd, P.weights, target_output_ix=target)


def _eval(acq, X):
    Xg = X.clone().requires_grad_(True)
    kg = acq(Xg.unsqueeze(1))
    (g,) = torch.autograd.grad(kg.sum(), Xg)
    return kg.detach().cpu().numpy(), g.cpu().numpy()


@pytest.mark.parametrize("case", [
    dict(target=0, kw=dict(n_train=30, n_disc=200, n_cand=12, n_train_per_obj=(30, 22))),          # single-CTA path
    dict(target=1, kw=dict(n_train=40, n_disc=1500, n_cand=40, n_scal=6, seed=3), dev=True),      # staged pipeline, CUDA tensors
    dict(target=None, kw=dict(n_train=24, n_disc=300, n_cand=10, n_scal=3, seed=5)),              # coupled evaluation
], ids=["small-decoupled-ragged", "staged-decoupled", "coupled"])
def test_fifty_appends_track_a_fresh_plan(case):
    P = small_problem(**case["kw"])
    dev = torch.device("cuda") if case.get("dev") else torch.device("cpu")
    X = P.candidates.to(dev)
    xd = P.x_disc.to(dev)
    acq = _fresh(P, case["target"], xd)
    _eval(acq, X)  # builds the plan
    plan0 = acq._plan
    rng = np.random.default_rng(11)
    worst = 0.0
    for step in range(50):
        m = int(rng.integers(0, 2)) if step % 3 else step % 2  # ragged: objectives grow at different rates
        x = torch.tensor(rng.random(P.d))
        y = float(rng.normal())
        assert acq.append_observation(m, x, y) is True
        assert acq._plan is plan0  # extended in place, not rebuilt
        if step % 10 == 9 or step == 0:
            kg, g = _eval(acq, X)
            assert acq._plan is plan0  # ... and the model fingerprint matches the extended plan
            ref = _fresh(P, case["target"], xd)  # P.model was extended by append_observation
            kg_r, g_r = _eval(ref, X)
            scale = float(ref._get_plan().read("A0").abs().max())
            np.testing.assert_allclose(kg, kg_r, rtol=1e-9, atol=1e-12 * scale)
            np.testing.assert_allclose(g, g_r, rtol=1e-6, atol=1e-10 * scale)
            worst = max(worst, float(np.max(np.abs(kg - kg_r)) / scale))
            # (posterior means at the discretisation: both mean caches carry ~cond(K) eps of rounding)
            np.testing.assert_allclose(acq._plan.read("mu_disc").cpu().numpy(), ref._get_plan().read("mu_disc").cpu().numpy(),
                                       rtol=1e-9, atol=1e-11 * scale)
            ref.invalidate()
    ns = [o.n for o in P.model.models]
    assert sum(ns) == sum(case["kw"].get("n_train_per_obj", (case["kw"]["n_train"],) * 2)) + 50 and ns[0] != ns[1]
    print(f"max |dKG| / max|intercept| over the checks: {worst:.2e}")


def test_append_without_room_falls_back_to_a_rebuild():
    P = small_problem(n_train=126, n_disc=150, n_cand=6)  # capacity 128 rows: two appends fit, the third does not
    acq = _fresh(P, 0)
    _eval(acq, P.candidates)
    rng = np.random.default_rng(3)
    results = []
    for _ in range(4):
        results.append(acq.append_observation(0, torch.tensor(rng.random(2)), float(rng.normal())))
    assert results[:2] == [True, True] and results[2] is False
    kg, g = _eval(acq, P.candidates)  # rebuilt from the (already extended) model at this call
    kg_r, g_r = _eval(_fresh(P, 0), P.candidates)
    assert np.array_equal(kg, kg_r) and np.array_equal(g, g_r)


def test_native_append_errors():
    from decoupledbo_b200 import _native

    P = small_problem()
    acq = _fresh(P, 0)
    _eval(acq, P.candidates)
    with pytest.raises(ValueError):
        acq._plan.append_point(5, torch.zeros(2), 0.0)  # objective out of range
    with pytest.raises(ValueError):
        acq._plan.append_point(0, torch.zeros(3), 0.0)  # wrong dimension
