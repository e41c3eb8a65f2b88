"""A few c4-shaped forward+backward passes of ONE objective (ncu target; developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

C = int(os.environ.get("C", 4096)); N = int(os.environ.get("N", 16384)); S = int(os.environ.get("S", 16))
tgt = int(os.environ.get("TGT", 0)); reps = int(os.environ.get("REPS", 3))
P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N)
dev = torch.device("cuda")
X = P.candidates.to(dev)
acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=None if tgt < 0 else tgt)
plan = acq._get_plan()
for _ in range(reps):
    kg, dX = plan.forward_device(X, True)
torch.cuda.synchronize()
print("ok", float(kg.sum()), plan.stats())
