import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from test_gpu_accuracy import _matern_ld, _chol_ld, _solve_ld, LD
from helpers import oracle_model
from oracle import discretekg as odk, gp as ogp
P = synthetic.problem_c2(n_cand=6)
tgt = 1
acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=tgt)
with torch.no_grad(): kg = acq(P.candidates.unsqueeze(1))
plan = acq._get_plan()
o = P.model.models[tgt]; x = o.train_x.numpy(); ls = o.lengthscale.numpy()
K = _matern_ld(x, x, ls, o.outputscale) + LD(o.noise) * np.eye(o.n, dtype=LD); L = _chol_ld(K)
B_ld = _solve_ld(L, _matern_ld(x, P.x_disc.numpy(), ls, o.outputscale))
Bg = plan.read("B").cpu().numpy()
print('B abs err', float(np.abs(Bg - B_ld).max()), 'B scale', float(np.abs(B_ld).max()))
om = oracle_model(P.model, distance="direct")
for c in (0, 4):
    xc = P.candidates[c].numpy(); pts = np.concatenate([xc[None], P.x_disc.numpy()])
    kx = _matern_ld(xc[None], x, ls, o.outputscale)[0]
    sol = _solve_ld(L, kx[:, None])[:, 0]
    cov = _matern_ld(xc[None], pts, ls, o.outputscale)[0] - _matern_ld(pts, x, ls, o.outputscale) @ sol
    var = cov[0] + LD(o.noise); z = cov / np.sqrt(var)
    sl = plan.read("slopes")[c].cpu().numpy(); vg = plan.read("var")[c].item()
    # oracle variants
    m_d, cov_d = ogp.posterior(om.models[tgt], torch.tensor(pts), False)
    _, row, vn = ogp.posterior_row(om.models[tgt], torch.tensor(xc), P.x_disc)
    # gpu-association in numpy fp64 with LAPACK-solved B: kx . B
    print(c, 'var rel err gpu %.2e | slopes abs err gpu %.2e | oracle dense cov-row err %.2e | oracle row err %.2e | cov scale %.2e' % (
        float(abs(vg - var) / var), float(np.abs(sl[:-1] - z[1:]).max()), float(np.abs(cov_d[0].numpy() - cov).max()), float(np.abs(row.numpy() - cov).max()), float(np.abs(cov).max())))
print('kg gpu', kg.numpy())
