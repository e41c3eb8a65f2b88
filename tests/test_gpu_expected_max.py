"""dkg_expected_max_lines_dev (C-ABI) against the reference's known answers
(test_discretekg.py:138-342), the reference-generated golden vectors, and the oracle."""
import math
import re

import numpy as np
import pytest
import torch

from helpers import load_golden
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def run(a, b, cap=64, grad=False):
    from decoupledbo_b200 import _native

    a = torch.as_tensor(a, dtype=torch.double).reshape(1, -1)
    b = torch.as_tensor(b, dtype=torch.double).reshape(1, -1)
    out = _native.expected_max_lines(a, b, hull_cap=cap, want_grad=grad)
    h = int(out["hull_count"][0])
    res = dict(e=float(out["emax"][0]), idx=out["hull_idx"][0, :h].cpu().tolist(),
               x=out["hull_x"][0, : max(h - 1, 0)].cpu().tolist(), h=h)
    if grad:
        res["da"] = out["dE_da"][0].cpu().numpy()
        res["db"] = out["dE_db"][0].cpu().numpy()
    return res


def test_empty_raises_value_error():  # test_discretekg.py:139-148, :264-277
    from decoupledbo_b200 import _native

    msg = "Expected inputs to specify at least one line. Got intercepts.shape[-1]=0."
    with pytest.raises(ValueError, match=re.escape(msg)):
        _native.expected_max_lines(torch.zeros(1, 0), torch.zeros(1, 0))


def test_reference_kats():
    r = run([1, 1.5], [0, 0])  # :150-158 zero-slope shortcut -> argmax only
    assert r["idx"] == [1] and r["x"] == [] and r["e"] == 1.5
    r = run([1.5], [-1.9])  # :160-167
    assert r["idx"] == [0] and r["e"] == pytest.approx(1.5)
    r = run([1.5, 0], [-0.5, 0])  # :169-182
    assert r["idx"] == [0, 1] and r["x"] == [3.0]
    r = run([0, 1.5], [0, -0.5])
    assert r["idx"] == [1, 0] and r["x"] == [3.0]
    r = run([0, 0, -0.5, 0], [-1, -1, 0, 1.5])  # :184-196 equal-slope regression
    assert r["idx"] == [0, 3] and r["x"] == [0.0]
    r = run([0, -1, 0], [-2, -1, 0])  # :198-215 dominated line
    assert r["idx"] == [0, 2] and r["x"] == [0.0]
    r = run([-1, 0, 0], [-1, 0, -2])
    assert r["idx"] == [2, 1] and r["x"] == [0.0]
    # expectation KATs (:279-328)
    assert run([1.5], [0])["e"] == pytest.approx(1.5)
    assert run([0], [1])["e"] == pytest.approx(0, abs=1e-300)
    assert run([0, 0], [0, 1])["e"] == pytest.approx(1 / math.sqrt(2 * math.pi))
    # (the reference's "hump" KAT :312-328 is not an upper envelope -- it is pinned on the oracle's
    #  expectation function in test_oracle_kats.py; its convex counterpart is E[1 + |Z|])
    assert run([0, 1, 1, 0], [0, 1, -1, 0])["e"] == pytest.approx(1 + math.sqrt(2 / math.pi))
    # tiny slopes (1e-12 < 1e-9 tolerance) take the shortcut (:363)
    assert run([1.5, 0], [0, 1e-12])["idx"] == [0]


def test_golden_line_sets():
    G = load_golden("epigraph_golden.npz")
    for k in range(int(G["n_sets"])):
        a, b = G[f"a{k}"], G[f"b{k}"]
        r = run(a, b, cap=max(64, len(a)), grad=True)
        e = float(G[f"e{k}"])
        assert abs(r["e"] - e) <= 1e-15 * max(1.0, np.abs(a).max()), (k, r["e"], e)  # erf/exp ulps
        np.testing.assert_allclose(r["x"], G[f"x{k}"], rtol=0, atol=0, err_msg=str(k))
        want_idx = G[f"idx{k}"].tolist()
        if r["idx"] != want_idx:  # only identical duplicate lines may be swapped
            assert len(r["idx"]) == len(want_idx)
            for i, j in zip(r["idx"], want_idx):
                assert a[i] == a[j] and b[i] == b[j], (k, r["idx"], want_idx)
        else:
            # dE/da = Phi differences, dE/db = -phi differences: CUDA's erf/exp and the host libm
            # differ by an ulp, so a difference of two of them may be off by a few 1e-16
            np.testing.assert_allclose(r["da"], G[f"ga{k}"], rtol=0, atol=1e-15)
            np.testing.assert_allclose(r["db"], G[f"gb{k}"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("L", [1, 2, 31, 33, 1000, 5000])
def test_batched_random_vs_oracle(L):
    from decoupledbo_b200 import _native

    rng = np.random.default_rng(L)
    P = 24
    a = rng.normal(size=(P, L))
    b = rng.normal(size=(P, L)) * rng.choice([1e-3, 1.0, 50.0], size=(P, 1))
    out = _native.expected_max_lines(torch.tensor(a), torch.tensor(b), hull_cap=64, want_grad=True)
    for p in range(P):
        E, idx, dp, dq, x = odk.expected_max_gradients_np(a[p], b[p])
        assert abs(float(out["emax"][p]) - E) <= 1e-15 * max(1.0, abs(E))
        h = int(out["hull_count"][p])
        assert h == len(idx)
        assert out["hull_idx"][p, : min(h, 64)].cpu().tolist() == idx[:64].tolist()
        da = np.zeros(L)
        db = np.zeros(L)
        da[idx] = dp
        db[idx] = dq
        np.testing.assert_allclose(out["dE_da"][p].cpu().numpy(), da, atol=1e-15)
        np.testing.assert_allclose(out["dE_db"][p].cpu().numpy(), db, atol=1e-15)


def test_every_line_on_the_hull_and_slow_path():
    """Tangents of a parabola: all L lines are hull vertices, so the chord filter keeps all of
    them (more than the survivor capacity) and the exact fallback path must be taken."""
    L = 3000
    b = np.linspace(-3, 3, L)
    a = -0.5 * b * b
    r = run(a, b, cap=L)
    E, idx, *_ = odk.expected_max_gradients_np(a, b)
    assert r["h"] == L == len(idx)
    assert r["idx"] == idx.tolist()
    assert abs(r["e"] - E) <= 1e-14
