"""Edge shapes through the C-ABI: ragged / tiny / odd sizes the tile-padded kernels must survive
(the reference's own edge cases are the empty and single-line sets, covered in
test_gpu_expected_max.py; these cover the GP side)."""
import numpy as np
import pytest
import torch

from helpers import oracle_model
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _problem(d, n_train, n_disc, n_scal, n_cand, n_obj, seed, kernel=0):
    from decoupledbo_b200 import synthetic

    ls = [0.3 + 0.2 * m for m in range(n_obj)]
    osc = [1.0 + m for m in range(n_obj)]
    P = synthetic.make_problem(
        "edge", d, n_train, ls, osc, [0.05] * n_obj, [1e-2] * n_obj,
        synthetic.sobol(n_disc, d, 50 + seed), n_scal, n_cand, seed_train=60 + seed, seed_cand=70 + seed,
        seed_w=seed)
    for o in P.model.models:
        o.kernel = kernel
    if n_obj != 2:  # simplex weights helper is for any n_obj, keep as is
        pass
    return P


CASES = [
    # d, n_train, N, S, C, M
    (1, 3, 1, 1, 1, 1),       # everything minimal: one training-side line + the candidate's own
    (1, 5, 2, 1, 3, 1),       # single objective (default weights [[1.]])
    (2, 17, 7, 3, 5, 2),      # nothing is a multiple of a tile
    (3, 33, 129, 2, 130, 2),  # N, C just above one tile
    (8, 20, 40, 4, 9, 2),     # maximum supported input dimension
    (2, 40, 300, 5, 31, 3),   # three objectives
    (2, 16, 1000, 17, 4, 2),  # S not a power of two
]


@pytest.mark.parametrize("case", CASES, ids=[f"d{c[0]}n{c[1]}N{c[2]}S{c[3]}C{c[4]}M{c[5]}" for c in CASES])
@pytest.mark.parametrize("kernel", [0, 1], ids=["matern", "rbf"])
def test_odd_shapes_match_oracle(case, kernel):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    d, n_train, N, S, C, M = case
    P = _problem(d, n_train, N, S, C, M, seed=sum(case), kernel=kernel)
    om = oracle_model(P.model)
    targets = list(range(M)) + [None]
    for target in targets:
        W = P.weights if M > 1 else None
        if M == 1 and target is None:
            acq = DiscreteKnowledgeGradient(P.model, P.x_disc)  # single output, default weights
            W_or = torch.ones(1, 1, dtype=torch.double)
        else:
            acq = DiscreteKnowledgeGradient(P.model, P.x_disc, W if W is not None else torch.ones(1, 1, dtype=torch.double),
                                            target_output_ix=target)
            W_or = W if W is not None else torch.ones(1, 1, dtype=torch.double)
        X = P.candidates.clone().requires_grad_(True)
        kg = acq(X.unsqueeze(1))
        (g,) = torch.autograd.grad(kg.sum(), X)
        Xo = P.candidates.clone().requires_grad_(True)
        want = odk.forward(om, Xo.unsqueeze(1), P.x_disc, W_or, target, dense=True)
        (go,) = torch.autograd.grad(want.sum(), Xo)
        scale = max(1.0, float(torch.cat([o.train_y for o in P.model.models]).abs().max()))
        np.testing.assert_allclose(kg.detach().numpy(), want.detach().numpy(), rtol=1e-9, atol=1e-12 * scale)
        np.testing.assert_allclose(g.numpy(), go.numpy(), rtol=1e-7, atol=1e-10 * scale)


def test_not_positive_definite_reports_cleanly():
    """Duplicate training points with zero noise: K is singular; GPyTorch's jitter retries
    (1e-8 .. 1e-6) rescue it, exactly singular + huge scale must raise instead of returning NaNs."""
    from decoupledbo_b200 import gp_state
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    x = torch.tensor([[0.2, 0.3], [0.2, 0.3], [0.7, 0.1]], dtype=torch.double)
    objs = [gp_state.GPObjective(train_x=x, train_y=torch.tensor([0.1, 0.1, -0.3]), lengthscale=torch.tensor([0.5, 0.5]),
                                 outputscale=1.0, mean_const=0.0, noise=0.0)]
    acq = DiscreteKnowledgeGradient(gp_state.GPModelList(objs), torch.rand(5, 2, dtype=torch.double))
    kg = acq(torch.rand(4, 1, 2, dtype=torch.double))  # jitter path
    assert torch.isfinite(kg).all()
    objs[0].outputscale = 1e12  # jitter of 1e-6 is far below the rounding of a 1e12-scaled matrix
    acq2 = DiscreteKnowledgeGradient(gp_state.GPModelList(objs), torch.rand(5, 2, dtype=torch.double))
    with pytest.raises(RuntimeError, match="positive definite"):
        acq2(torch.rand(4, 1, 2, dtype=torch.double))
