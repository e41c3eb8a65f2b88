"""Time of one incremental append (dkg_plan_append_point) against a fresh plan at the c4 shape (developer tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import numpy as np, torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient

for (name, P) in (("c4 (n=400, N=16384, S=16)", synthetic.problem_c4(n_cand=256)),
                  ("c2 (n=100, N=1024, S=16)", synthetic.problem_c2(n_cand=64)),
                  ("BO preset (n=60, N=121)", synthetic.make_problem("bo", 2, 60, [0.2, 1.8], [1.0, 50.0], [0.0, 0.0], [1e-4, 1e-4],
                                                                      synthetic.std_grid(11, 2), 16, 8, seed_train=5, seed_cand=6))):
    dev = torch.device("cuda")
    xd = P.x_disc.to(dev); X = P.candidates.to(dev)
    for tgt in (0, 1):
        t_build = []
        for _ in range(4):
            acq = DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=tgt)
            torch.cuda.synchronize(); t0 = time.perf_counter(); acq._get_plan(); torch.cuda.synchronize()
            t_build.append(time.perf_counter() - t0)
        rng = np.random.default_rng(1)
        t_app = {"target": [], "other": []}
        for k in range(16):
            m = k % 2
            x = torch.tensor(rng.random(P.d)); y = float(rng.normal())
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ok = acq.append_observation(m, x, y)
            torch.cuda.synchronize()
            assert ok
            t_app["target" if m == tgt else "other"].append(time.perf_counter() - t0)
        with torch.no_grad():
            a = acq(X.unsqueeze(1))
            b = DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=tgt)(X.unsqueeze(1))
        err = float((a - b).abs().max())
        print(f"{name} target {tgt}: fresh plan {sorted(t_build[1:])[1]*1e3:.2f} ms | append to the fantasised objective "
              f"{np.median(t_app['target'])*1e3:.2f} ms, to the other objective {np.median(t_app['other'])*1e3:.2f} ms | "
              f"max |KG(appended) - KG(fresh)| after 16 appends {err:.2e}")
