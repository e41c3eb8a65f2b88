"""Live comparison with the reference's own functions (only where /root/reference is mounted,
i.e. the build container; skipped on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import discretekg as odk
from oracle import reference_loader as rl

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted")


def test_random_line_sets_bit_exact():
    ref = rl.load_reference_discretekg()
    rng = np.random.default_rng(99)
    for k in range(300):
        n = int(rng.integers(1, 150))
        a = torch.tensor(rng.normal(size=n))
        b = torch.tensor(rng.normal(size=n))
        if k % 3 == 0:
            b = torch.round(b * 2) / 2
        if k % 7 == 0:
            b = b * 1e-10
        i1, x1 = ref.calculate_epigraph_indices(a, b)
        i2, x2 = odk.epigraph_indices(a, b)
        assert torch.equal(i1, i2) and torch.equal(x1, x2)
        e1 = ref.calculate_expected_value_of_piecewise_linear_function(a[i1], b[i1], x1)
        e2 = odk.expected_value_of_piecewise_linear_function(a[i2], b[i2], x2)
        assert e1.item() == e2.item()


def test_std_grid_matches_reference():
    rl.load_reference_discretekg()
    from decoupledbo.modules.utils import make_torch_std_grid as ref_grid

    from decoupledbo_b200.modules.utils import make_torch_std_grid

    for npa, d in ((3, 2), (4, 3), (11, 2), (2, 4)):
        assert torch.equal(ref_grid(npa, d, {"dtype": torch.double}), make_torch_std_grid(npa, d, {"dtype": torch.double}))
        assert torch.equal(ref_grid(npa, d, {"dtype": torch.double}), odk.make_std_grid(npa, d))
