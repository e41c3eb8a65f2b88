import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/decoupled-kg_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from helpers import oracle_model, small_problem
from oracle import discretekg as odk, gp as ogp
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
torch.set_default_dtype(torch.double)
P = small_problem()
om = oracle_model(P.model)
for tgt in (0,1):
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc, P.weights, target_output_ix=tgt)
    X = P.candidates.clone().requires_grad_(True)
    kg = acq(X.unsqueeze(1)); kg.sum().backward()
    Xo = P.candidates.clone().requires_grad_(True)
    want = odk.forward(om, Xo.unsqueeze(1), P.x_disc, P.weights, tgt, dense=True); want.sum().backward()
    print('tgt',tgt,'kg', kg[:4].tolist(), 'want', want[:4].tolist())
    print(' max abs err', (kg-want).abs().max().item(), 'max rel', ((kg-want).abs()/want.abs().clamp_min(1e-300)).max().item())
    print(' grad err', (X.grad-Xo.grad).abs().max().item(), 'grad scale', Xo.grad.abs().max().item())
    plan = acq._get_plan()
    # stage checks
    a,b = odk.lines_single_output(om, P.candidates[0], tgt, P.x_disc, P.weights, dense=True)
    sl = plan.read('slopes').cpu(); A0 = plan.read('A0').cpu(); an = plan.read('a_new').cpu()
    z = b[0]/P.weights[0,tgt]
    print(' slope err', (sl[0,:-1]-z[1:]).abs().max().item(), (sl[0,-1]-z[0]).abs().item(), 'A0 err', (A0[0]-a[0,1:]).abs().max().item(), 'anew err', (an[0,0]-a[0,0]).abs().item())
    print(' stats', plan.stats())
