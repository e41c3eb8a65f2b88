"""Oracle against the committed golden vectors (made by the REFERENCE's own functions in the build
container: oracle/make_golden.py)."""
import numpy as np
import torch

from helpers import GOLDEN_KG_SPECS, load_golden, oracle_model, small_problem
from oracle import discretekg as odk


def test_epigraph_and_expectation_bit_exact():
    G = load_golden("epigraph_golden.npz")
    for k in range(int(G["n_sets"])):
        a = torch.tensor(G[f"a{k}"], requires_grad=True)
        b = torch.tensor(G[f"b{k}"], requires_grad=True)
        idx, x = odk.epigraph_indices(a, b)
        assert np.array_equal(idx.numpy(), G[f"idx{k}"]), k
        assert np.array_equal(x.detach().numpy(), G[f"x{k}"]), k
        e = odk.expected_value_of_piecewise_linear_function(a[idx], b[idx], x)
        assert e.item() == float(G[f"e{k}"]), k
        ga, gb = torch.autograd.grad(e, (a, b), allow_unused=True)
        ga = np.zeros(len(a)) if ga is None else ga.numpy()
        gb = np.zeros(len(b)) if gb is None else gb.numpy()
        np.testing.assert_allclose(ga, G[f"ga{k}"], rtol=0, atol=1e-15)
        np.testing.assert_allclose(gb, G[f"gb{k}"], rtol=0, atol=1e-15)


def test_kg_assembly_matches_reference_code():
    """Oracle forward (dense and row-only) == the reference's forward code run on the same GP
    posterior (pins discretekg.py:131-338 above the GP boundary, decoupled AND coupled)."""
    G = load_golden("kg_reference_code_golden.npz")
    for name, spec in GOLDEN_KG_SPECS.items():
        P = small_problem(**spec)
        om = oracle_model(P.model)
        for key, target in (("t0", 0), ("t1", 1), ("coupled", None)):
            want = G[f"{name}__{key}__kg"]
            X = P.candidates.clone().requires_grad_(True)
            got = odk.forward(om, X.unsqueeze(1), P.x_disc, P.weights, target, dense=True)
            np.testing.assert_allclose(got.detach().numpy(), want, rtol=1e-12, atol=1e-15)
            (g,) = torch.autograd.grad(got.sum(), X)
            np.testing.assert_allclose(g.numpy(), G[f"{name}__{key}__grad"], rtol=1e-9, atol=1e-13)
            got_row = odk.forward(om, P.candidates.unsqueeze(1), P.x_disc, P.weights, target, dense=False)
            np.testing.assert_allclose(got_row.numpy(), want, rtol=1e-9, atol=1e-13)
