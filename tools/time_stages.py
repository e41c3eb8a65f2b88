"""Per-stage CUDA-event times of c4-shaped forward+backward passes, both objectives (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

C = int(os.environ.get("C", 4096)); N = int(os.environ.get("N", 16384)); S = int(os.environ.get("S", 16))
P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N)
dev = torch.device("cuda")
X = P.candidates.to(dev)
for tgt in (0, 1):
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=tgt)
    plan = acq._get_plan()
    for _ in range(3):
        plan.forward_device(X, True)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    for _ in range(5):
        plan.forward_device(X, True)
    prof = _native.profile_read(); _native.profile_enable(False)
    print(f"tgt {tgt}: " + "  ".join(f"{k}={v[0]/5:.3f}" for k, v in prof.items() if v[1]) + f"  total={sum(v[0] for v in prof.values())/5:.3f}  stats={plan.stats()[:4]}")
