"""Timing of the COUPLED evaluation (target_output_ix=None, reference calculate_discrete_kg,
discretekg.py:162-235) at the c2 and c4 shapes: one forward(+backward) over all candidates, CUDA events."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

dev = torch.device("cuda")
for name, P in (("c2", synthetic.problem_c2(n_cand=512)), ("c4", synthetic.problem_c4(n_cand=int(os.environ.get("C", 4096))))):
    X = P.candidates.to(dev)
    S = int(P.weights.shape[0])
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=None)
    plan = acq._get_plan()
    for grad in (False, True):
        for _ in range(2):
            plan.forward_device(X, grad)
        torch.cuda.synchronize()
        _native.launch_count_reset()
        _native.profile_enable(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            kg, dX = plan.forward_device(X, grad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        prof = _native.profile_read(); _native.profile_enable(False)
        print("   " + "  ".join(f"{k}={v[0]/3:.3f}ms/{v[1]//3}" for k, v in prof.items() if v[1]))
        print(json.dumps({"shape": name, "C": X.shape[0], "N": P.x_disc.shape[0], "S": S, "grad": grad, "ms": round(ms, 3),
                          "kg_evals_per_s": X.shape[0] * S / ms * 1e3, "launches": _native.launch_count() // 3,
                          "stats": plan.stats()}))
    acq.invalidate()
