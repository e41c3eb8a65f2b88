"""Shared helpers for the parity tests (oracle <-> CUDA path)."""
import numpy as np
import torch

from decoupledbo_b200 import synthetic
from oracle import gp as ogp


def oracle_model(model, distance="gpytorch"):
    return ogp.OracleModelList(
        [ogp.OracleObjective(**kw, distance=distance) for kw in synthetic.to_oracle_kwargs(model)]
    )


def rel_err(got, want, floor=0.0):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), floor)


def small_problem(d=2, n_train=30, n_disc=64, n_scal=4, n_cand=16, seed=0, noise=(1e-2, 1e-3),
                  lengthscales=(0.3, 0.6), outputscales=(1.0, 2.5), kernel=0, y_std=(1.0, 1.0),
                  y_mean=(0.0, 0.0), n_train_per_obj=None):
    P = synthetic.make_problem(
        "small", d, n_train, list(lengthscales), list(outputscales), [0.1, 0.0], list(noise),
        synthetic.sobol(n_disc, d, 100 + seed), n_scal, n_cand, seed_train=200 + seed,
        seed_cand=300 + seed, seed_w=seed,
    )
    for m, o in enumerate(P.model.models):
        o.kernel = kernel
        o.y_std = float(y_std[m])
        o.y_mean = float(y_mean[m])
        if n_train_per_obj is not None:
            k = n_train_per_obj[m]
            o.train_x = o.train_x[:k].clone()
            o.train_y = o.train_y[:k].clone()
    return P


# problems behind tests/golden/kg_reference_code_golden.npz (oracle/make_golden.py)
GOLDEN_KG_SPECS = {
    "matern_d2": dict(d=2, n_train=30, n_disc=64, n_scal=4, n_cand=10, kernel=0, seed=0),
    "rbf_d3": dict(d=3, n_train=25, n_disc=50, n_scal=3, n_cand=8, kernel=1, seed=1),
    "std_d2": dict(d=2, n_train=20, n_disc=36, n_scal=5, n_cand=8, kernel=0, seed=2,
                   y_std=(2.5, 0.5), y_mean=(1.0, -3.0)),
    "ragged_d2": dict(d=2, n_train=24, n_disc=49, n_scal=2, n_cand=6, kernel=0, seed=3,
                      n_train_per_obj=(24, 17)),
}


def load_golden(name):
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def refit_reference_problem():
    """The reference's own test fixture (tests/modules/acquisition/conftest.py:30-47: 10 Sobol
    points, randn targets, MAP-fitted SingleTaskGPs) re-derived by oracle/refit_reference_fixture.py,
    together with the reference's golden KG values for it (test_discretekg.py:62,78,93,107)."""
    from decoupledbo_b200.gp_state import GPModelList, GPObjective

    G = load_golden("reference_fixture_refit.npz")
    objs = [GPObjective(train_x=torch.tensor(G["train_x"]), train_y=torch.tensor(G["train_y"][:, m]),
                        lengthscale=torch.tensor(G["lengthscale"][m]), outputscale=float(G["outputscale"][m]),
                        mean_const=float(G["mean_const"][m]), noise=float(G["noise"][m])) for m in range(2)]
    n = 3
    disc = torch.stack([torch.repeat_interleave(torch.linspace(0, 1, n), n),
                        torch.tile(torch.linspace(0, 1, n), (n,))]).T.double()  # test_discretekg.py:17-25
    W = torch.tensor([[0.7, 0.3], [0.6, 0.4], [0.5, 0.5]], dtype=torch.double)  # conftest.py:60-66
    target_x = torch.tensor([[[[0.5, 0.5]], [[0, 1]], [[0, 0.5]]], [[[0, 0]], [[1, 0]], [[0.5, 0]]]],
                            dtype=torch.double)  # test_discretekg.py:30-46
    return GPModelList(objs), disc, W, target_x, G
