"""Latency of one public-API call at the sizes the reference's BO loop runs (developer tool):
plan build, first call (graph capture), steady-state distribution; native call vs full Python path."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

torch.set_default_dtype(torch.double)
def pct(v, q): v = sorted(v); return v[min(len(v) - 1, int(q * len(v)))]
for (grid, ntr, C) in ((11, 60, 10), (11, 60, 1), (32, 100, 64), (32, 100, 512)):
    P = synthetic.make_problem("bo", 2, ntr, [0.2, 1.8], [1.0, 50.0], [0.0, 0.0], [1e-4, 1e-4],
                               synthetic.std_grid(grid, 2), 16, 8, seed_train=5, seed_cand=6)
    X = torch.rand(C, 2)
    t0 = time.perf_counter()
    acq = DiscreteKnowledgeGradient(P.model, synthetic.std_grid(grid, 2), P.weights, target_output_ix=0)
    plan = acq._get_plan(); torch.cuda.synchronize()
    t_plan = time.perf_counter() - t0
    for grad in (False, True):
        t0 = time.perf_counter(); plan.forward_host(X, grad); t_first = time.perf_counter() - t0
        nat = []
        for _ in range(300):
            t0 = time.perf_counter(); plan.forward_host(X, grad); nat.append(time.perf_counter() - t0)
        full = []
        for _ in range(300):
            t0 = time.perf_counter()
            x = X.unsqueeze(1).clone().requires_grad_(grad); v = acq(x)
            if grad: torch.autograd.grad(v.sum(), x)
            full.append(time.perf_counter() - t0)
        # device-only time of the same call (events around forward_device on resident inputs)
        Xd = X.cuda(); plan.forward_device(Xd, grad); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): plan.forward_device(Xd, grad)
        e1.record(); torch.cuda.synchronize()
        print(f"N={grid*grid} n={ntr} C={C} grad={grad}: plan {t_plan*1e3:.1f} ms, first call {t_first*1e6:.0f} us | native host call "
              f"median {statistics.median(nat)*1e6:.0f} p99 {pct(nat,0.99)*1e6:.0f} max {max(nat)*1e6:.0f} us | "
              f"public API median {statistics.median(full)*1e6:.0f} p99 {pct(full,0.99)*1e6:.0f} us | device-side {e0.elapsed_time(e1)/50*1e3:.0f} us/call")
