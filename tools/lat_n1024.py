"""Device-side latency of one forward (+ backward) at the c2-like shape N = 1024, n = 100 (developer tool; A/B switches via env)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

torch.set_default_dtype(torch.double)
out = []
for C in (64, 512):
    P = synthetic.make_problem("bo", 2, 100, [0.2, 1.8], [1.0, 50.0], [0.0, 0.0], [1e-4, 1e-4],
                               synthetic.std_grid(32, 2), 16, 8, seed_train=5, seed_cand=6)
    acq = DiscreteKnowledgeGradient(P.model, synthetic.std_grid(32, 2), P.weights, target_output_ix=0)
    plan = acq._get_plan()
    Xd = torch.rand(C, 2).cuda()
    for grad in (False, True):
        for _ in range(5): plan.forward_device(Xd, grad)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): plan.forward_device(Xd, grad)
        e1.record(); torch.cuda.synchronize()
        out.append(f"C={C} grad={int(grad)}: {e0.elapsed_time(e1)/200*1e3:.0f} us")
    if os.environ.get("PROF"):
        _native.profile_enable(True)
        for _ in range(20): plan.forward_device(Xd, True)
        prof = _native.profile_read(); _native.profile_enable(False)
        out.append("   " + "  ".join(f"{k}={v[0]/20*1e3:.0f}" for k, v in prof.items() if v[1]))
print(" | ".join(out))
