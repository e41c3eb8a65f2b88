"""DiscreteKgOptimisationSpec (strategy.py:166-273 mirror) driving the CUDA path end to end."""
import numpy as np
import pytest
import torch

from helpers import oracle_model, small_problem
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def test_optimize_for_single_objective_batched_restarts():
    from decoupledbo_b200 import _native
    from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec

    P = small_problem(n_train=12, noise=(1e-2, 1e-2))
    om = oracle_model(P.model)
    spec = DiscreteKgOptimisationSpec(n_discretisation_points_per_axis=5, num_restarts=6,
                                      raw_samples=32, batch_limit=6, max_iter=40)
    torch.manual_seed(0)
    _native.launch_count_reset()
    x, i, v = spec.optimize_for_single_objective(P.model, [1.0, 1.0], 2, scalarisation_weights=P.weights)
    assert _native.launch_count() > 0
    assert x.shape == (1, 2) and i in (0, 1)
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    grid = odk.make_std_grid(5, 2)
    want = odk.kg_single_output(om, x[0], i, grid, P.weights, dense=True)
    np.testing.assert_allclose(float(v), float(want), rtol=1e-8, atol=1e-12)
    # the optimiser must not end below the best raw Sobol sample of the chosen objective
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    acq = DiscreteKnowledgeGradient(P.model, grid, P.weights, target_output_ix=i)
    with torch.no_grad():
        raw = acq(torch.rand(64, 1, 2, dtype=torch.double))
    assert float(v) >= 0.0
    assert float(v) >= 0.5 * float(raw.max())


def test_reference_batch_limit_one_gives_same_api():
    """batch_limit=1 (the reference preset, bo_loop.py:127-129) still works: one row per call."""
    from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec

    P = small_problem(n_train=10, noise=(1e-2, 1e-2))
    spec = DiscreteKgOptimisationSpec(3, num_restarts=2, raw_samples=4, batch_limit=1, max_iter=5)
    torch.manual_seed(1)
    x, i, v = spec.optimize_for_single_objective(P.model, torch.tensor([1.0, 2.0]), 2,
                                                  scalarisation_weights=P.weights)
    assert x.shape == (1, 2) and isinstance(i, int) and torch.is_tensor(v)


def test_optimize_for_full_evaluation_coupled():
    """strategy.py:242-273: coupled evaluation (target_output_ix=None) through optimize_acqf."""
    from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec

    P = small_problem(n_train=12, noise=(1e-2, 1e-2))
    om = oracle_model(P.model)
    spec = DiscreteKgOptimisationSpec(4, num_restarts=4, raw_samples=16, batch_limit=4, max_iter=30)
    torch.manual_seed(3)
    x, v = spec.optimize_for_full_evaluation(P.model, 2, scalarisation_weights=P.weights)
    assert x.shape == (1, 2) and float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    want = odk.kg_coupled(om, x[0], odk.make_std_grid(4, 2), P.weights, dense=True)
    np.testing.assert_allclose(float(v), float(want), rtol=1e-8, atol=1e-12)


def test_concurrent_objectives_give_the_serial_result():
    """f1: the per-objective optimisations run side by side (threads + streams) but the starting points are
    drawn in objective order, so the chosen (x, objective, value) equals the serial loop's bit for bit."""
    from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec

    P = small_problem(n_train=14, noise=(1e-2, 1e-2))
    out = {}
    for concurrent in (True, False):
        spec = DiscreteKgOptimisationSpec(5, num_restarts=6, raw_samples=24, batch_limit=6, max_iter=25)
        spec.concurrent_objectives = concurrent
        torch.manual_seed(7)
        out[concurrent] = spec.optimize_for_single_objective(P.model, [1.0, 1.5], 2, scalarisation_weights=P.weights)
    (xa, ia, va), (xb, ib, vb) = out[True], out[False]
    assert ia == ib and torch.equal(xa, xb) and torch.equal(va, vb)


def test_evaluate_objectives_matches_the_individual_acquisition_functions():
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
    from decoupledbo_b200.multi import evaluate_objectives

    P = small_problem(n_disc=400, n_cand=37)
    for dev in ("cpu", "cuda"):
        X = P.candidates.to(dev)
        acqs = [DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=i) for i in range(2)]
        kg, dX = evaluate_objectives(acqs, X, need_grad=True)
        assert kg.shape == (2, 37) and dX.shape == (2, 37, 2) and kg.device.type == dev
        for i, a in enumerate(acqs):
            Xg = X.clone().requires_grad_(True)
            v = a(Xg.unsqueeze(1))
            (g,) = torch.autograd.grad(v.sum(), Xg)
            assert torch.equal(kg[i], v.detach()) and torch.equal(dX[i], g)
        kg2, none = evaluate_objectives(acqs, X, need_grad=False)
        assert none is None and torch.equal(kg2, kg)
