"""BASELINE.json configs[2] ("c3"): the c2 problem (d=2, n_train=100, |X_disc|=1024, 16 scalarisations)
with forward+backward driven through optimize_acqf -- 64 restarts x 512 raw samples, all restarts of
an objective in one batched call per L-BFGS iteration (batch_limit = num_restarts) -- via
DiscreteKgOptimisationSpec.optimize_for_single_objective, CPU float64 tensors in and out (the
reference's TKWARGS).  Reports wall time, acquisition calls, KG evaluations and evals/s."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from decoupledbo_b200.modules.acquisition_optimisation_strategy import DiscreteKgOptimisationSpec

torch.set_default_dtype(torch.double)
P = synthetic.problem_c2(n_cand=8)
S = int(P.weights.shape[0])

# count what the optimiser asks of the acquisition function (candidates per call, with / without grad)
calls = {"fwd": 0, "fwd_bwd": 0, "cand_fwd": 0, "cand_fwd_bwd": 0}
_orig = DiscreteKnowledgeGradient.forward
def _counting(self, X):
    n = X.numel() // X.shape[-1]
    if X.requires_grad:
        calls["fwd_bwd"] += 1; calls["cand_fwd_bwd"] += n
    else:
        calls["fwd"] += 1; calls["cand_fwd"] += n
    return _orig(self, X)
DiscreteKnowledgeGradient.forward = _counting

# host / GPU split: time spent inside the native call (H2D + kernels + D2H + sync) vs everything else
native_s = {"t": 0.0}
_fh = _native.Plan.forward_host
def _timed_fh(self, *a, **k):
    t0 = time.perf_counter()
    try:
        return _fh(self, *a, **k)
    finally:
        native_s["t"] += time.perf_counter() - t0
_native.Plan.forward_host = _timed_fh

for concurrent in (True, False):
    spec = DiscreteKgOptimisationSpec(32, num_restarts=64, raw_samples=512, batch_limit=64, max_iter=200)
    spec.concurrent_objectives = concurrent
    rows = []
    for rep in range(6):  # SAME seed every time: the spread is the machine's, not the optimiser's path
        torch.manual_seed(0)
        for k in calls: calls[k] = 0
        native_s["t"] = 0.0
        _native.launch_count_reset()
        t0 = time.perf_counter()
        x, i, v = spec.optimize_for_single_objective(P.model, [1.0, 1.0], 2, scalarisation_weights=P.weights)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        evals = (calls["cand_fwd"] + calls["cand_fwd_bwd"]) * S
        rows.append({"concurrent_objectives": concurrent, "seconds": dt, "seconds_in_native_calls": native_s["t"],
                     "calls_fwd": calls["fwd"], "calls_fwd_bwd": calls["fwd_bwd"],
                     "kg_evals": evals, "kg_evals_per_s": evals / dt, "launches": _native.launch_count(),
                     "chosen_objective": int(i), "value_per_cost": float(v)})
        print(json.dumps(rows[-1]))
    warm = sorted(r["seconds"] for r in rows[1:])
    print(json.dumps({"config": "c3: c2 problem, optimize_acqf 64 restarts x 512 raw samples, maxiter 200, both objectives",
                      "concurrent_objectives": concurrent, "median_seconds": warm[len(warm) // 2],
                      "min_seconds": warm[0], "max_seconds": warm[-1],
                      "spread": (warm[-1] - warm[0]) / warm[len(warm) // 2],
                      "median_kg_evals_per_s": sorted(r["kg_evals_per_s"] for r in rows[1:])[(len(rows) - 1) // 2]}))
