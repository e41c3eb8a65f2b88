"""Wall time of one BO-iteration's acquisition optimisation in the reference's production
configuration (bo_loop.py:123-131: 11x11 grid, 16 scalarisations, 10 restarts, 32 raw samples,
maxiter 200; both objectives; the reference pins batch_limit=1, here batch_limit=num_restarts).
The reference publishes this number (10-13 s median per iteration on a CPU cluster)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

torch.set_default_dtype(torch.double)
n_train = int(os.environ.get("NTR", 60))
P = synthetic.make_problem("bo", 2, n_train, [0.2, 1.8], [1.0, 50.0], [0.0, 0.0], [1e-4, 1e-4],
                           synthetic.std_grid(11, 2), 16, 8, seed_train=5, seed_cand=6)
# per-call latency of the public API at the production batch size (10 restarts)
acq = DiscreteKnowledgeGradient(P.model, synthetic.std_grid(11, 2), P.weights, target_output_ix=0)
X = torch.rand(10, 1, 2)
for need_grad in (False, True):
    for _ in range(5):
        x = X.clone().requires_grad_(need_grad); v = acq(x)
        if need_grad: torch.autograd.grad(v.sum(), x)
    t0 = time.perf_counter()
    for _ in range(200):
        x = X.clone().requires_grad_(need_grad); v = acq(x)
        if need_grad: torch.autograd.grad(v.sum(), x)
    print(f"forward{'+backward' if need_grad else ''} call, C=10, N=121, n={n_train}: {(time.perf_counter()-t0)/200*1e6:.0f} us")
for bl, label in ((10, "batch_limit=num_restarts"), (1, "batch_limit=1 (reference preset)")):
    spec = DiscreteKgOptimisationSpec(11, num_restarts=10, raw_samples=32, batch_limit=bl, max_iter=200)
    ts = []
    for rep in range(4):
        torch.manual_seed(rep)
        _native.launch_count_reset()
        t0 = time.perf_counter()
        x, i, v = spec.optimize_for_single_objective(P.model, [1.0, 1.0], 2, scalarisation_weights=P.weights)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"config": label, "n_train": n_train, "seconds_per_bo_iteration_acqf_optimisation": sorted(ts)[len(ts)//2],
                      "all": ts, "kernel_launches": _native.launch_count(), "chosen_objective": i, "value": float(v)}))
