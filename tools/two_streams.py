import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200"))
import torch
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
P = synthetic.problem_c4()
dev = torch.device("cuda"); X = P.candidates.to(dev); xd = P.x_disc.to(dev)
plans = [DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=i)._get_plan() for i in (0, 1)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def seq():
    for p in plans: p.forward_device(X, True)
def par():
    cur = torch.cuda.current_stream()
    for p, s in zip(plans, streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s): p.forward_device(X, True)
    for s in streams: cur.wait_stream(s)
for name, fn in (("sequential", seq), ("two streams", par)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms/step")
