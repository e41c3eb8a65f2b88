"""Host-side logic of the drop-in (no GPU): constructor validation and error types, GP-state
extraction from a duck-typed ModelListGP, grid builder, objective choice, and that the product
fails loudly (instead of falling back) without a CUDA device."""
import types

import pytest
import torch

from decoupledbo_b200 import _native, gp_state, synthetic
from decoupledbo_b200.botorch_compat import BotorchTensorDimensionError, UnsupportedError
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from decoupledbo_b200.modules.acquisition_optimisation_strategy import (
    DiscreteKgOptimisationSpec,
    choose_best_objective,
)
from decoupledbo_b200.modules.utils import make_torch_std_grid
from helpers import small_problem


@pytest.fixture()
def problem():
    return small_problem()


def test_constructor_validation(problem):  # discretekg.py:92-119
    P = problem
    with pytest.raises(BotorchTensorDimensionError):
        DiscreteKnowledgeGradient(P.model, P.x_disc[0], P.weights, 0)
    with pytest.raises(UnsupportedError):
        DiscreteKnowledgeGradient(P.model, P.x_disc, None, 0)
    with pytest.raises(BotorchTensorDimensionError):
        DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights[0], 0)
    with pytest.raises(BotorchTensorDimensionError):
        DiscreteKnowledgeGradient(P.model, P.x_disc, torch.ones(3, 5), 0)
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, 1)
    assert acq.target_output_ix == 1 and acq.x_discretisation is P.x_disc
    assert acq.scalarisation_weights is P.weights and acq.model is P.model
    with pytest.raises(UnsupportedError):  # discretekg.py:125-129
        acq.set_X_pending(None)


def test_single_output_default_weights():
    P = small_problem()
    single = gp_state.GPModelList(P.model.models[:1])
    acq = DiscreteKnowledgeGradient(single, P.x_disc, None, 0)
    assert acq.scalarisation_weights.tolist() == [[1.0]]  # discretekg.py:105


def test_forward_shape_errors(problem):
    P = problem
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, 0)
    with pytest.raises(RuntimeError, match="last dimension"):  # discretekg.py:137-141
        acq(torch.zeros(3, 1, P.d + 1))
    with pytest.raises(AssertionError):  # t_batch_mode_transform(expected_q=1)
        acq(torch.zeros(3, 2, P.d))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(problem):
    P = problem
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, 0)
    with pytest.raises(_native.NativeLibraryError, match="no CPU fallback"):
        acq(P.candidates.unsqueeze(1))


def _duck_single(o, with_transform):
    ns = types.SimpleNamespace
    gp = ns(
        train_inputs=(o.train_x,), train_targets=o.train_y,
        covar_module=ns(outputscale=torch.tensor(o.outputscale),
                        base_kernel=type("MaternKernel", (), {})()),
        mean_module=ns(constant=torch.tensor(o.mean_const)),
        likelihood=ns(noise=torch.tensor([o.noise])),
    )
    gp.covar_module.base_kernel.lengthscale = o.lengthscale.reshape(1, -1)
    gp.covar_module.base_kernel.nu = 2.5
    if with_transform:
        gp.outcome_transform = ns(means=torch.tensor([[1.5]]), stdvs=torch.tensor([[2.0]]))
    return gp


def test_extract_state_from_duck_typed_model_list(problem):
    P = problem
    duck = types.SimpleNamespace(models=[_duck_single(o, m == 1) for m, o in enumerate(P.model.models)], num_outputs=2)
    st = gp_state.extract_gp_state(duck)
    assert st.num_outputs == 2
    for m, (a, b) in enumerate(zip(st.models, P.model.models)):
        assert torch.equal(a.train_x, b.train_x) and torch.equal(a.train_y, b.train_y)
        assert torch.equal(a.lengthscale, b.lengthscale)
        assert a.outputscale == b.outputscale and a.noise == b.noise and a.kernel == gp_state.MATERN52
        assert (a.y_mean, a.y_std) == ((1.5, 2.0) if m == 1 else (0.0, 1.0))
    with pytest.raises(TypeError):
        gp_state.extract_gp_state(object())


def test_problem_blob_constructor():
    sd = {}
    for m in range(2):
        sd[f"models.{m}.likelihood.noise_covar.raw_noise"] = torch.tensor([-float("inf")])
        sd[f"models.{m}.likelihood.noise_covar.raw_noise_constraint.lower_bound"] = torch.tensor(1e-8)
    blob = dict(train_x=torch.rand(7, 2), train_y=torch.rand(7, 2), model_state_dict=sd,
                fixed_hyperparams=dict(length_scales=[0.2, 1.8], output_scales=[1, 50], means=[0, 0]))
    ml = gp_state.model_from_problem_blob(blob)
    assert ml.models[1].outputscale == 50.0 and ml.models[0].lengthscale.tolist() == [0.2, 0.2]
    assert ml.models[0].noise == pytest.approx(1e-8)


def test_std_grid_order():  # utils.py:83-92 docstring example
    g = make_torch_std_grid(3, 2)
    want = [[0, 0], [0, .5], [0, 1], [.5, 0], [.5, .5], [.5, 1], [1, 0], [1, .5], [1, 1]]
    assert g.tolist() == want
    with pytest.raises(ValueError):
        make_torch_std_grid(3, 0)


def test_choose_best_objective_tie_breaking():  # strategy.py:143-163
    t = torch.tensor
    x0, x1 = t([[0.1]]), t([[0.2]])
    assert choose_best_objective([(0, x0, t(0.2)), (1, x1, t(0.3))], [1.0, 1.0])[0] == 1
    assert choose_best_objective([(0, x0, t(0.2)), (1, x1, t(0.3))], [1.0, 2.0])[0] == 0  # per cost
    # both negative -> clipped to 0 -> the cheaper objective wins; returned value is NOT clipped
    i, x, v = choose_best_objective([(0, x0, t(-0.2)), (1, x1, t(-0.1))], [2.0, 1.0])
    assert i == 1 and float(v) == pytest.approx(-0.1)
    # exact tie and equal costs -> first in objective order
    assert choose_best_objective([(0, x0, t(0.5)), (1, x1, t(0.5))], [1.0, 1.0])[0] == 0


def test_spec_surface():
    spec = DiscreteKgOptimisationSpec(n_discretisation_points_per_axis=3, num_restarts=2,
                                      raw_samples=4, batch_limit=1, max_iter=200)
    assert spec._options() == {"batch_limit": 1, "maxiter": 200}
    assert spec._discretisation(2).shape == (9, 2)


def test_synthetic_problem_shapes():
    P = synthetic.problem_c2(n_cand=8)
    assert P.x_disc.shape == (1024, 2) and P.weights.shape == (16, 2)
    assert torch.allclose(P.weights.sum(-1), torch.ones(16, dtype=torch.double))
    P2 = synthetic.problem_c2(n_cand=8)
    assert torch.equal(P.candidates, P2.candidates) and torch.equal(P.model.models[1].train_y, P2.model.models[1].train_y)


def test_precision_attribute_is_an_addition_not_a_signature_change(problem):
    """The constructor keeps the reference signature (discretekg.py:62-68); the reduced-precision mode is
    an attribute, validated on the host, and switching it drops the cached native state."""
    import inspect

    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    params = list(inspect.signature(DiscreteKnowledgeGradient.__init__).parameters)
    assert params == ["self", "model", "x_discretisation", "scalarisation_weights", "target_output_ix"]
    P = problem
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=0)
    assert acq.precision == "float64"
    acq.precision = "float32"
    assert acq.precision == "float32" and acq._plan is None
    with pytest.raises(ValueError):
        acq.precision = "tf32"
