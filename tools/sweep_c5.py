"""c5 sweep (BASELINE.json configs[4]): |X_disc| x scalarisations x candidates on one GPU.
Prints one line per shape: ms per forward+backward (one objective), KG evals/s, and a spot parity
check of a few candidates against the oracle (row-only posterior)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from helpers import oracle_model
from oracle import discretekg as odk

shapes = [(1024, 16, 512), (4096, 16, 4096), (16384, 16, 4096), (16384, 64, 4096), (16384, 256, 2048),
          (65536, 16, 4096), (65536, 64, 2048), (16384, 16, 32768)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
dev = torch.device("cuda")
out = []
for (N, S, C) in shapes:
    P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N)
    om = oracle_model(P.model)
    X = P.candidates.to(dev)
    res = {"N": N, "S": S, "C": C}
    for tgt in (0, 1):
        acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=tgt)
        plan = acq._get_plan()
        for _ in range(2):
            kg, dX = plan.forward_device(X, True)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            kg, dX = plan.forward_device(X, True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        spots = [0, C // 2, C - 1]
        errs, gerrs = [], []
        scale = float(plan.read("A0").abs().max())
        for c in spots:
            x = P.candidates[c].clone().requires_grad_(True)
            v = odk.kg_single_output(om, x, tgt, P.x_disc, P.weights, dense=False); v.backward()
            errs.append(abs(float(kg[c]) - v.item()) / max(abs(v.item()), 1e-3 * scale * 1e-9 + 1e-300) if abs(v.item()) > 1e-12 * scale else abs(float(kg[c]) - v.item()) / scale)
            gerrs.append(float((dX[c].cpu() - x.grad).abs().max()) / max(float(x.grad.abs().max()), 1e-12 * scale))
        res[f"ms_t{tgt}"] = round(ms, 3); res[f"evals_per_s_t{tgt}"] = C * S / ms * 1e3
        res[f"max_rel_err_t{tgt}"] = max(errs); res[f"max_grad_rel_err_t{tgt}"] = max(gerrs); res[f"stats_t{tgt}"] = plan.stats()
        acq.invalidate()
    print(json.dumps(res), flush=True)
