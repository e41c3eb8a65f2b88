"""ORACLE support: generate tests/golden/*.npz with the REFERENCE's own functions.

Run in the build container only (needs /root/reference; botorch names come from oracle/_stubs):

    python oracle/make_golden.py

Writes
  tests/golden/epigraph_golden.npz  -- random / degenerate line sets with the outputs of the
      reference's calculate_epigraph_indices and
      calculate_expected_value_of_piecewise_linear_function (discretekg.py:341-452), plus the
      gradients autograd gives for the expectation of the upper envelope.
  tests/golden/kg_reference_code_golden.npz -- KG values produced by the reference's
      calculate_discrete_kg_conditioning_on_single_output / calculate_discrete_kg /
      DiscreteKnowledgeGradient.forward (discretekg.py:131-338) running on top of the ORACLE GP
      posterior (oracle/gp.py through the stub ModelListGP): pins everything above the GP
      boundary; the GP posterior itself is third-party code absent from this container.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))

from oracle import reference_loader as rl  # noqa: E402


def line_sets(rng):
    sets = []
    for t in range(160):
        n = int(rng.integers(1, 400)) if t % 4 else int(rng.integers(1, 12))
        a = rng.normal(size=n)
        b = rng.normal(size=n)
        if t % 3 == 0:
            b = np.round(b * 2) / 2  # many equal slopes
        if t % 5 == 0:
            a = np.round(a * 2) / 2  # ties in the intercepts (and identical lines with t % 15 == 0)
        if t % 7 == 0:
            b = b * 1e-10  # all |slopes| < 1e-9 -> shortcut branch
        if t % 11 == 0:
            b = np.sort(b)
            a = -0.5 * b * b  # every line is a hull vertex (tangents of a parabola)
        if t % 13 == 0:
            b = b * 1e-3 + 5.0  # hull far from z = 0
        sets.append((a, b))
    return sets


def main():
    torch.set_default_dtype(torch.double)
    ref = rl.load_reference_discretekg()
    rng = np.random.default_rng(20261018)
    out = {}
    sets = line_sets(rng)
    out["n_sets"] = np.array(len(sets))
    for k, (a, b) in enumerate(sets):
        ta = torch.tensor(a, requires_grad=True)
        tb = torch.tensor(b, requires_grad=True)
        idx, inter = ref.calculate_epigraph_indices(ta, tb)
        e = ref.calculate_expected_value_of_piecewise_linear_function(ta[idx], tb[idx], inter)
        ga, gb = torch.autograd.grad(e, (ta, tb), allow_unused=True)
        out[f"a{k}"] = a
        out[f"b{k}"] = b
        out[f"idx{k}"] = idx.numpy().astype(np.int64)
        out[f"x{k}"] = inter.detach().numpy()
        out[f"e{k}"] = np.array(e.item())
        out[f"ga{k}"] = np.zeros_like(a) if ga is None else ga.numpy()
        out[f"gb{k}"] = np.zeros_like(b) if gb is None else gb.numpy()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "epigraph_golden.npz"), **out)
    print("wrote epigraph_golden.npz with", len(sets), "line sets")

    # ---- KG through the reference's own code on top of the oracle GP ----
    from decoupledbo_b200 import synthetic
    from oracle import gp as ogp

    cases = {}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import GOLDEN_KG_SPECS, small_problem  # noqa: E402

    for name, sp in GOLDEN_KG_SPECS.items():
        P = small_problem(**sp)
        om = ogp.OracleModelList(
            [ogp.OracleObjective(**kw) for kw in synthetic.to_oracle_kwargs(P.model)])
        rm = rl.wrap_model_for_reference(om)
        res = {}
        for target in (0, 1, None):
            acq = ref.DiscreteKnowledgeGradient(rm, P.x_disc, P.weights, target_output_ix=target)
            X = P.candidates.clone().requires_grad_(True)
            kg = acq(X.unsqueeze(1))
            (g,) = torch.autograd.grad(kg.sum(), X)
            key = "coupled" if target is None else f"t{target}"
            res[key] = (kg.detach().numpy(), g.numpy())
        for key, (kg, g) in res.items():
            out_k = f"{name}__{key}"
            cases[out_k + "__kg"] = kg
            cases[out_k + "__grad"] = g
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kg_reference_code_golden.npz"), **cases)
    print("wrote kg_reference_code_golden.npz:", sorted(cases)[:4], "...")


if __name__ == "__main__":
    main()
