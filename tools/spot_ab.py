"""A/B of the accuracy-relevant switches on three c4 spot candidates against the row-only oracle (developer tool)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from decoupledbo_b200 import synthetic
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
    P = synthetic.problem_c4(n_cand=4096)
    dev = torch.device("cuda")
    acq = DiscreteKnowledgeGradient(P.model, P.x_disc.to(dev), P.weights, target_output_ix=0)
    plan = acq._get_plan()
    kg, dX = plan.forward_device(P.candidates.to(dev), True)
    spots = [0, 2048, 4095, 1, 2, 3]
    print(json.dumps({"kg": [float(kg[c]) for c in spots], "scale": float(plan.read("A0").abs().max())}))
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from decoupledbo_b200 import synthetic
from helpers import oracle_model
from oracle import discretekg as odk
P = synthetic.problem_c4(n_cand=4096)
om = oracle_model(P.model)
spots = [0, 2048, 4095, 1, 2, 3]
want = [odk.kg_single_output(om, P.candidates[c].clone(), 0, P.x_disc, P.weights, dense=False).item() for c in spots]
print("oracle", want)
for name, env in (("default", {}), ("track", {"DKG_ZSTAT_TRACK": "1"}), ("ng8", {"DKG_OZ_DIAGONALS": "8"}), ("tdmma", {"DKG_T_GEMM": "dmma"}),
                  ("sample", {"DKG_PROBE_SAMPLE": "1"}), ("f64filter", {"DKG_FILTER": "f64"}), ("noshort", {"DKG_HULL_SHORT": "0"})):
    out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(name, ["%.3e" % (abs(a - b)) for a, b in zip(d["kg"], want)], "scale %.3g" % d["scale"])
    except Exception:
        print(name, "FAILED", out.stderr[-500:])
