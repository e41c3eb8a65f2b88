import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from helpers import oracle_model
from oracle import discretekg as odk
N, S, C = 16384, 16, 4096
P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N)
om = oracle_model(P.model); dev = torch.device("cuda"); X = P.candidates.to(dev)
for tgt in (0, 1):
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=tgt)
    plan = acq._get_plan(); kg, dX = plan.forward_device(X, True)
    scale = float(plan.read("A0").abs().max())
    for c in (0, C // 2, C - 1):
        v = odk.kg_single_output(om, P.candidates[c], tgt, P.x_disc, P.weights, dense=False).item()
        print(os.environ.get("DKG_T_SOLVE"), os.environ.get("DKG_COV_GEMM"), "tgt", tgt, "c", c, "kg %.6e oracle %.6e abs %.2e rel %.2e (scale %.2e)" % (float(kg[c]), v, abs(float(kg[c]) - v), abs(float(kg[c]) - v) / max(abs(v), 1e-300), scale))
