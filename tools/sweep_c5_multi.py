"""BASELINE.json configs[4] ("c5"): |X_disc| x scalarisations x candidates on ALL ranks of a torchrun launch
(the candidates of one problem are split across the GPUs by evaluate_objectives; both objectives, forward +
backward), with the vectorised CPU baseline (oracle/vectorised.py, rank 0, bounded sample) beside each shape.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_c5_multi.py
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "decoupled-kg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
group = None
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
from decoupledbo_b200 import synthetic
from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
from decoupledbo_b200.multi import evaluate_objectives

shapes = [(1024, 16, 512), (4096, 16, 4096), (16384, 16, 4096), (16384, 64, 4096), (16384, 256, 2048),
          (65536, 16, 4096), (65536, 64, 2048), (16384, 16, 32768)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
cpu = os.environ.get("CPU_BASELINE", "1") == "1"
for (N, S, C) in shapes:
    P = synthetic.problem_c4(n_cand=C, n_scal=S, n_disc=N)
    xd = P.x_disc.to(dev)
    acqs = [DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=i) for i in range(2)]
    X = P.candidates.to(dev)
    for _ in range(2):
        kg, dX = evaluate_objectives(acqs, X, need_grad=True, group=group)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        kg, dX = evaluate_objectives(acqs, X, need_grad=True, group=group)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.double, device=dev)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res = {"N": N, "S": S, "C": C, "gpus": world, "ms_both_objectives_fwd_bwd": round(float(ms), 3),
           "kg_evals_per_s": 2 * C * S / float(ms) * 1e3, "argmax": [int(v) for v in kg.argmax(dim=1)]}
    if rank == 0 and cpu:
        from helpers import oracle_model
        from oracle import vectorised as ov
        torch.set_num_threads(os.cpu_count() or 1)
        om = oracle_model(P.model)
        prep = ov.Prepared(om, P.x_disc, P.weights, 0)
        nb = 8
        t0 = time.perf_counter()
        kg_cpu, _ = ov.kg_batch(prep, P.candidates[:nb], need_grad=True)
        dt = time.perf_counter() - t0
        res["cpu_vectorised_evals_per_s"] = nb * S / dt
        res["cpu_cores"] = os.cpu_count()
        res["cpu_sample"] = f"{nb} candidates x objective 0 x {S} scalarisations, fwd+bwd ({dt:.1f} s)"
        res["max_abs_err_vs_cpu_on_sample"] = float((kg[0, :nb].cpu() - kg_cpu).abs().max())
        res["gpu_over_cpu"] = res["kg_evals_per_s"] / res["cpu_vectorised_evals_per_s"]
    if rank == 0:
        print(json.dumps(res), flush=True)
    for a in acqs: a.invalidate()
    del acqs, X, xd
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
