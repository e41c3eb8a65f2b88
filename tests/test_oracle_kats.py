"""The oracle's hull / expectation stage against every known-answer test the reference holds for
it: /root/reference/tests/modules/acquisition/test_discretekg.py:138-342 (13 epigraph cases, 6
expectation cases).  Values are the reference's; the test bodies are ours."""
import math
import re

import pytest
import torch

from oracle import discretekg as odk

t = torch.tensor


def close(a, b, **kw):
    torch.testing.assert_close(a, b, **kw)


class TestEpigraphKATs:
    def test_empty_raises_value_error(self):  # test_discretekg.py:139-148
        msg = "Expected inputs to specify at least one line. Got intercepts.shape[-1]=0."
        with pytest.raises(ValueError, match=re.escape(msg)):
            odk.epigraph_indices(t([]), t([]))

    def test_all_zero_slopes_shortcut(self):  # :150-158
        idx, x = odk.epigraph_indices(t([1, 1.5]), t([0.0, 0.0]))
        close(idx, t([1]))
        assert x.numel() == 0 and x.dtype == torch.double

    def test_single_line(self):  # :160-167
        idx, x = odk.epigraph_indices(t([1.5]), t([-1.9]))
        close(idx, t([0]))
        assert x.numel() == 0

    @pytest.mark.parametrize("ordered", [True, False])
    def test_two_lines(self, ordered):  # :169-182
        a, b = t([1.5, 0.0]), t([-0.5, 0.0])
        if not ordered:
            a, b = a.flip(0), b.flip(0)
        idx, x = odk.epigraph_indices(a, b)
        close(idx, t([0, 1] if ordered else [1, 0]))
        close(x, t([3.0]))

    def test_two_equal_slopes_regression(self):  # :184-196
        idx, x = odk.epigraph_indices(t([0, 0, -0.5, 0]), t([-1, -1, 0, 1.5]))
        close(idx, t([0, 3]))
        close(x, t([0.0]))

    @pytest.mark.parametrize("order,expected", [([0, 1, 2], [0, 2]), ([1, 2, 0], [2, 1])])
    def test_dominated_line_ignored(self, order, expected):  # :198-215
        a, b = t([0.0, -1.0, 0.0])[order], t([-2.0, -1.0, 0.0])[order]
        idx, x = odk.epigraph_indices(a, b)
        close(idx, t(expected))
        close(x, t([0.0]))

    @pytest.mark.parametrize("slopes", [[-0.5, 0], [0, 1e-12], [-0.5, -0.5]])
    def test_intersection_gradcheck(self, slopes):  # :217-235
        a = t([1.5, 0.0], requires_grad=True)
        b = t([float(s) for s in slopes], requires_grad=True)
        torch.autograd.gradcheck(lambda *args: odk.epigraph_indices(*args)[1], (a, b), raise_exception=True)

    @pytest.mark.parametrize("offset", [0, 1])
    def test_gradients_two_of_four_slopes_identical(self, offset):  # :237-260
        a = t([offset, offset, -0.5, 0.0], requires_grad=True)
        b = t([-1.0, -1.0, 0.0, 1.5], requires_grad=True)
        _, x = odk.epigraph_indices(a, b)
        only = x.squeeze(0)
        assert only.ndim == 0
        (gb,) = torch.autograd.grad(only, b, retain_graph=True)
        (ga,) = torch.autograd.grad(only, a, retain_graph=True)
        close(gb, t([0.16 * offset, 0.0, 0.0, -0.16 * offset]))
        close(ga, t([0.4, 0.0, 0.0, -0.4]))


class TestExpectationKATs:
    def test_empty_raises(self):  # :264-277
        msg = "Expected inputs to specify at least one line. Got intercepts.shape[-1]=0."
        with pytest.raises(ValueError, match=re.escape(msg)):
            odk.expected_value_of_piecewise_linear_function(t([]), t([]), t([]))

    def test_constant(self):  # :279-288
        assert odk.expected_value_of_piecewise_linear_function(t([1.5]), t([0.0]), t([])) == pytest.approx(1.5)

    def test_sloped_line(self):  # :290-299
        assert odk.expected_value_of_piecewise_linear_function(t([0.0]), t([1.0]), t([])) == pytest.approx(0)

    def test_relu(self):  # :301-310
        v = odk.expected_value_of_piecewise_linear_function(t([0.0, 0.0]), t([0.0, 1.0]), t([0.0]))
        assert v == pytest.approx(1 / math.sqrt(2 * math.pi))

    def test_hump(self):  # :312-328
        v = odk.expected_value_of_piecewise_linear_function(
            t([0.0, 1.0, 1.0, 0.0]), t([0.0, 1.0, -1.0, 0.0]), t([-1.0, 0.0, 1.0]))
        want = math.erf(1 / math.sqrt(2)) - (1 - math.exp(-1 / 2)) * math.sqrt(2 / math.pi)
        assert v == pytest.approx(want)

    def test_gradcheck(self):  # :330-342
        a = t([0.0, 1.0, 1.0, 0.0], requires_grad=True)
        b = t([0.0, 1.0, -1.0, 0.0], requires_grad=True)
        x = t([-1.0, 0.0, 1.0], requires_grad=True)
        torch.autograd.gradcheck(odk.expected_value_of_piecewise_linear_function, (a, b, x), raise_exception=True)

    def test_wrong_boundary_shape(self):  # discretekg.py:425-429
        with pytest.raises(odk.OracleTensorDimensionError):
            odk.expected_value_of_piecewise_linear_function(t([0.0, 1.0]), t([0.0, 1.0]), t([0.0, 1.0]))


def test_closed_form_gradient_matches_autograd():
    """Envelope theorem: dE/da_k = dPhi_k, dE/db_k = -dphi_k on the hull."""
    import numpy as np

    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(2, 80))
        a = torch.tensor(rng.normal(size=n), requires_grad=True)
        b = torch.tensor(rng.normal(size=n), requires_grad=True)
        e = odk.expected_max_of_lines(a, b)
        ga, gb = torch.autograd.grad(e, (a, b))
        E, idx, p, q, _ = odk.expected_max_gradients_np(a.detach().numpy(), b.detach().numpy())
        da = np.zeros(n)
        db = np.zeros(n)
        da[idx] = p
        db[idx] = q
        assert abs(E - e.item()) < 1e-14
        np.testing.assert_allclose(ga.numpy(), da, atol=1e-14)
        np.testing.assert_allclose(gb.numpy(), db, atol=1e-14)
        assert abs(da.sum() - 1.0) < 1e-14
