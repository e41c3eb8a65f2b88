"""The backward merges the hull records of a candidate in batches of what its shared-memory tables hold (2560
records; only S >= 64 on noisy objectives ever needs a second batch).  DKG_FIN_BATCH shrinks the batch so that the
multi-batch path runs at test sizes: the gradients must agree with the single-batch result up to summation order,
and with the oracle -- decoupled (finalize_kernel) and coupled (finalize_coupled_kernel)."""
import os

import numpy as np
import pytest
import torch

from helpers import oracle_model
from oracle import discretekg as odk

pytestmark = pytest.mark.gpu


def _grad(P, target, X, batch):
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

    old = os.environ.get("DKG_FIN_BATCH")
    os.environ["DKG_SMALL"] = "0"  # the staged pipeline's backward, whatever the size
    if batch is None:
        os.environ.pop("DKG_FIN_BATCH", None)
    else:
        os.environ["DKG_FIN_BATCH"] = str(batch)
    try:
        acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(X.device), P.weights, target_output_ix=target)
        Xg = X.clone().requires_grad_(True)
        kg = acq(Xg.unsqueeze(1))
        (g,) = torch.autograd.grad(kg.sum(), Xg)
        torch.cuda.synchronize()
        return kg.detach().cpu(), g.cpu()
    finally:
        os.environ.pop("DKG_SMALL", None)
        if old is None:
            os.environ.pop("DKG_FIN_BATCH", None)
        else:
            os.environ["DKG_FIN_BATCH"] = old


@pytest.mark.parametrize("target", [0, 1, None], ids=["decoupled0", "decoupled1", "coupled"])
def test_small_record_batches_give_the_same_gradient(target):
    from decoupledbo_b200 import synthetic

    P = synthetic.make_problem("fin", 2, 40, [0.25, 0.6], [1.0, 2.0], [0.05, 0.0], [1e-2, 1e-3],
                               synthetic.std_grid(24, 2), 12, 40, seed_train=21, seed_cand=22, seed_w=5)
    X = P.candidates.to("cuda")
    kg1, g1 = _grad(P, target, X, None)
    for batch in (7, 16, 33):  # several batches per candidate (12 scalarisations x ~5 hull records)
        kgb, gb = _grad(P, target, X, batch)
        assert torch.equal(kg1, kgb)
        scale = float(g1.abs().max())
        np.testing.assert_allclose(gb.numpy(), g1.numpy(), rtol=1e-10, atol=1e-12 * scale)
    if target is not None:
        om = oracle_model(P.model)
        for c in (0, 17, 39):
            x = P.candidates[c].clone().requires_grad_(True)
            v = odk.kg_single_output(om, x, target, P.x_disc, P.weights, dense=False)
            v.backward()
            np.testing.assert_allclose(_grad(P, target, X, 7)[1][c].numpy(), x.grad.numpy(), rtol=1e-6, atol=1e-9 * float(g1.abs().max()))
