"""Small odd-shaped evaluations that exercise every kernel variant (tiled statistics pass, fp32 filter in
both phases and in the row-mod layout, multi-level hull refinement, overflow, backward), meant to
run under `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_probe.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
os.environ["DKG_FILTER"] = "f32"  # the float filter also for batches below its size threshold
import torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

shapes = [  # d, n_train, N, S, C
    (2, 17, 7, 3, 5), (3, 33, 129, 2, 130), (2, 40, 4101, 5, 37), (4, 50, 4099, 16, 300), (2, 30, 1030, 17, 1025),
]
for (d, n, N, S, C) in shapes:
    P = synthetic.make_problem("san", d, n, [0.3, 0.5], [1.0, 2.0], [0.5, 0.05], [0.3, 1e-3],
                               synthetic.sobol(N, d, 5), S, C, seed_train=6, seed_cand=7, seed_w=1)
    for target in (0, 1, None):
        for prec in ("float64", "float32"):
            acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=target)
            acq.precision = prec
            X = P.candidates.clone().requires_grad_(True)
            kg = acq(X.unsqueeze(1))
            (g,) = torch.autograd.grad(kg.sum(), X)
            assert torch.isfinite(kg).all() and torch.isfinite(g).all()
            acq.invalidate()
    print("ok", (d, n, N, S, C), flush=True)
# every line on the hull (slow paths of the hull / overflow kernels) through the generic entry point
from decoupledbo_b200 import _native
z = torch.linspace(-3, 3, 3000, dtype=torch.double)
a = -(z * z)
out = _native.expected_max_lines(a.unsqueeze(0).repeat(3, 1), z.unsqueeze(0).repeat(3, 1)) if hasattr(_native, "expected_max_lines") else None
print("done")
