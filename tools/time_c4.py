"""Per-kernel timing of one c4-shaped forward(+backward) (developer tool; not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic, _native
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

C = int(os.environ.get("C", 4096)); N = int(os.environ.get("N", 16384)); S = int(os.environ.get("S", 16))
ntr = int(os.environ.get("NTR", 400))
t0 = time.time(); P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N, n_train=ntr); print("gen %.1fs" % (time.time() - t0))
dev = torch.device("cuda")
X = P.candidates.to(dev)
for tgt in (0, 1):
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=tgt)
    acq.precision = os.environ.get("PREC", "float64")
    torch.cuda.synchronize(); t0 = time.time(); plan = acq._get_plan(); torch.cuda.synchronize(); print("plan %.1f ms" % ((time.time() - t0) * 1e3))
    for grad in (False, True):
        for rep in range(3):
            plan.forward_device(X, grad)
        torch.cuda.synchronize()
        _native.profile_enable(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(3):
            kg, dX = plan.forward_device(X, grad)
        e1.record(); torch.cuda.synchronize()
        prof = _native.profile_read(); _native.profile_enable(False)
        tot = e0.elapsed_time(e1) / 3
        print(f"target {tgt} grad={grad}: {tot:.3f} ms/forward -> {C*S/tot*1e3:.3e} evals/s ; stats {plan.stats()}")
        print("   " + "  ".join(f"{k}={v[0]/3:.3f}ms/{v[1]//3}" for k, v in prof.items()))
    print("   kg range", float(kg.min()), float(kg.max()), "argmax", int(kg.argmax()))
