#!/usr/bin/env python
"""bench.py -- headline benchmark of the discrete-KG hot path on B200 (contract in the task).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU implementation

Metric (BASELINE.json): KG evals/s, 1 eval = one (candidate, objective, scalarisation) triple,
forward + backward.  Workload = BASELINE.json configs[3] ("c4", the configuration the target is
quoted on: 2-obj d=4 GP, n_train=400, |X_disc|=16384, 16 scalarisations, 4096 candidates), which
fits one GPU.  A "step" is one pass of DiscreteKnowledgeGradient forward+backward over the batch
of candidates for BOTH objectives (two acquisition functions, target_output_ix = 0 and 1), as
DiscreteKgOptimisationSpec.optimize_for_single_objective drives it, through the product API
`decoupledbo_b200.multi.evaluate_objectives` (both objectives on one upload of the candidates, one
CUDA stream per objective, one packed result buffer).  Multi-GPU: ONE problem, the global batch of
N x 4096 candidates (weak scaling) is split into contiguous shards with replicated GP state and the
values + gradients of BOTH objectives are all-gathered with ONE NCCL collective on the device, inside
the timed step; a "strong" block (4096 candidates in total over the N GPUs) is reported next to it.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "decoupled-kg_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "kg_evals_per_sec_fwd_bwd"
UNIT = "KG evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2"])
    ap.add_argument("--candidates", type=int, default=None, help="candidates per GPU (weak) / in total (strong)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU gets its own full batch (default); strong: one batch split over the GPUs")
    ap.add_argument("--precision", default="float64", choices=["float64", "float32"],
                    help="float32: reduced-precision mode (4-digit covariance contraction; not the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def build_problem(workload: str, n_cand: int):
    from decoupledbo_b200 import synthetic

    if workload == "c4":
        return synthetic.problem_c4(n_cand=n_cand)
    return synthetic.problem_c2(n_cand=n_cand)


def workload_config(workload: str, P, n_cand: int, world: int):
    o = P.model.models[0]
    return {
        "workload": f"{workload}: synthetic 2-obj d={P.d} GP, n_train={o.n}, |X_disc|={P.x_disc.shape[0]}, "
                    f"{P.weights.shape[0]} scalarisations, {n_cand} candidates/GPU, both objectives, fwd+bwd",
        "candidates_per_gpu": n_cand,
        "objectives": P.model.num_outputs,
        "scalarisations": int(P.weights.shape[0]),
        "x_disc": int(P.x_disc.shape[0]),
        "n_train": o.n,
        "d": P.d,
        "parallelism": f"candidate-sharded x{world} (one problem, contiguous shards of the global batch), replicated "
                       "GP state, ONE all-gather of [objectives, values+grads] on the device per step; the two "
                       "objectives' evaluations run on two CUDA streams",
        "l2": "no flush: per-step working set (slope rows 0.5 GB/objective + B, B^T 105 MB) exceeds the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------
# clocks (sampled DURING the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------
# CPU legs (the ONLY places bench.py touches oracle/)
# ------------------------------------------------------------------------------------------
def _oracle_model(P):
    from decoupledbo_b200 import synthetic
    from oracle import gp as ogp

    return ogp.OracleModelList(
        [ogp.OracleObjective(**kw) for kw in synthetic.to_oracle_kwargs(P.model)]
    )


def _host_mem_gb():
    try:
        import psutil

        return psutil.virtual_memory().available / 2**30
    except Exception:
        return 0.0


def cpu_reference_step(P, om, cand_ix: int, target: int, dense: bool):
    """One candidate x one objective x all scalarisations, forward + backward, reference-faithful
    CPU path (per-candidate evaluation, dense (N+1)^2 posterior covariance when `dense`, Python
    hull loop per scalarisation).  Returns seconds."""
    from oracle import discretekg as odk

    x = P.candidates[cand_ix].clone().requires_grad_(True)
    t0 = time.perf_counter()
    kg = odk.kg_single_output(om, x, target, P.x_disc, P.weights, dense=dense)
    kg.backward()
    return time.perf_counter() - t0


def run_reference_arm(args, rank: int, world: int):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores.
    botorch/gpytorch are not installable in this image, so the reference arm is the oracle port
    (same per-candidate structure, dense posterior covariance, Python hull loop)."""
    if rank != 0:
        return
    n_cand = args.candidates or (4096 if args.workload == "c4" else 512)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P = build_problem(args.workload, max(8, min(n_cand, 64)))
    om = _oracle_model(P)
    S = int(P.weights.shape[0])
    dense = _host_mem_gb() > 48.0 or P.x_disc.shape[0] <= 4096
    t_first = cpu_reference_step(P, om, 0, 0, dense)
    warm = args.warmup if t_first < 5.0 else min(args.warmup, 1)
    for w in range(1, warm):
        cpu_reference_step(P, om, w % P.candidates.shape[0], w % 2, dense)
    times = []
    for k in range(args.steps):
        times.append(cpu_reference_step(P, om, (k + warm) % P.candidates.shape[0], k % 2, dense))
    total = sum(times)
    value = S * args.steps / total
    sample = (f"each step = 1 candidate x 1 objective x {S} scalarisations of the same workload "
              f"(fwd+bwd, {'dense (N+1)^2 posterior covariance' if dense else 'row-only posterior'}, "
              f"Python hull loop); candidates are independent loop iterations (discretekg.py:145) so "
              f"evals/s extrapolates linearly")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, P, n_cand, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference repo is pure Python on botorch/gpytorch (absent here): oracle port timed",
        "warmup_requested": args.warmup,
        "warmup_note": ("warm-up cut to 1 step: one step of this arm costs > 5 s of CPU time (it only warms caches; "
                        "every step does the same work)") if warm != args.warmup else None,
        "gpus_note": "CPU arm: rank 0 alone runs it whatever --gpus says (task contract); it does not scale with N",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def measure_dgemm_peak(dev):
    """cuBLAS DGEMM 8192^3, best of 5 (burst) -- the fp64-tensor roofline denominator;
    MEASURED_PEAKS.json only carries HBM GB/s and bf16 TF/s."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.double, device=dev)
    b = torch.randn(n, n, dtype=torch.double, device=dev)
    for _ in range(2):
        a @ b
    best = float("inf")
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0 * n**3 / best / 1e9  # TFLOP/s


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from decoupledbo_b200 import _native
    from decoupledbo_b200.distributed import shard_bounds
    from decoupledbo_b200.modules.acquisition.discretekg import DiscreteKnowledgeGradient
    from decoupledbo_b200.multi import evaluate_objectives

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL's version banner / debug lines go to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    per_gpu = args.candidates or (4096 if args.workload == "c4" else 512)
    # weak scaling (the headline `value`): the global batch grows with the GPU count; strong: it is fixed
    n_weak = per_gpu * world
    n_strong = per_gpu
    n_main = n_weak if args.scaling == "weak" else n_strong
    P = build_problem(args.workload, n_weak)  # one problem; the strong batch is its first n_strong candidates
    S = int(P.weights.shape[0])
    M = P.model.num_outputs
    d = P.d
    xd = P.x_disc.to(dev)
    acqs = [DiscreteKnowledgeGradient(P.model, xd, P.weights, target_output_ix=i) for i in range(M)]
    for a in acqs:
        a.precision = args.precision
    plans = [a._get_plan() for a in acqs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        """W >= 3 warm-ups, then `steps` steps bracketed by barrier + synchronize; CUDA events on the
        launching stream, max over ranks.  Returns ms per step."""
        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        _native.launch_count_reset()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        timed.launches = _native.launch_count()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.double, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    def timed_wall(step, steps):
        for _ in range(3):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.double, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return 1e3 * float(t) / steps

    # ---- device-resident inputs: the global candidate batch lives in HBM on every rank ----
    def make_device_step(n_global):
        X_dev = P.candidates[:n_global].to(dev).contiguous()

        def step():
            kg, dX = evaluate_objectives(acqs, X_dev, need_grad=True, group=group)
            return kg.argmax(dim=1)  # first-index argmax per objective, identical on every rank

        return step

    # ---- host buffers: H2D of this rank's shard, kernels, ONE collective, ONE D2H, host argmax ----
    def make_host_step(n_global):
        X_host = P.candidates[:n_global].clone().pin_memory()

        def step():
            kg, dX = evaluate_objectives(acqs, X_host, need_grad=True, group=group)
            # (numpy: torch's CPU argmax costs ~4 ns per element single-threaded, 0.3 ms per step at 8 GPUs)
            return [int(np.argmax(kg[m].numpy())) for m in range(M)]

        return step

    step_main = make_device_step(n_main)
    for _ in range(2):
        step_main()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(step_main, args.steps)
    launches = timed.launches
    clocks = sampler.stop() if rank == 0 else None
    evals_step = n_main * M * S
    value = evals_step / (ms_step * 1e-3)
    lo, hi = shard_bounds(n_main, world, rank)
    n_cand = hi - lo  # candidates this rank evaluates per step

    e2e = None
    if not args.no_e2e:
        ms_e2e = timed_wall(make_host_step(n_main), args.steps)
        rows = -(-n_main // world)
        e2e = {
            "value": evals_step / (ms_e2e * 1e-3),
            "unit": UNIT,
            "h2d_bytes_per_step": rows * d * 8,
            "d2h_bytes_per_step": world * M * rows * (1 + d) * 8,
            "ms_per_step": ms_e2e,
            "api": "decoupledbo_b200.multi.evaluate_objectives(acqfs, X_host, need_grad=True, group): pinned host "
                   "candidates -> H2D of this rank's shard -> both objectives' kernels (dkg_forward_dev, one stream "
                   "each) -> one NCCL all-gather on the device (N > 1) -> one D2H of every rank's values + gradients "
                   "-> host argmax; bytes are per rank",
        }

    strong = None
    if world > 1 and args.scaling == "weak":
        ms_s = timed(make_device_step(n_strong), args.steps)
        strong = {"candidates_total": n_strong, "ms_per_step": ms_s, "value": n_strong * M * S / (ms_s * 1e-3),
                  "unit": UNIT}
        if not args.no_e2e:
            ms_se = timed_wall(make_host_step(n_strong), args.steps)
            strong["e2e_ms_per_step"] = ms_se
            strong["e2e_value"] = n_strong * M * S / (ms_se * 1e-3)

    # ---- roofline of the dominant kernel (separate profiled steps, CUDA events per launch) ----
    roofline = None
    cpu_baseline = None
    extra = {}
    if rank == 0:
        X_loc = P.candidates[lo:hi].to(dev).contiguous()
        _native.profile_enable(True)
        for _ in range(3):
            for plan in plans:
                plan.forward_device(X_loc, True)
        prof = _native.profile_read()
        _native.profile_enable(False)
        ms_g, n_g = prof["gemm_cov"]
        small_path = n_g == 0  # small discretisations: one fused CTA-per-candidate kernel (timed as "hull")
        if small_path:
            ms_g, n_g = prof["hull"]
        o = P.model.models[0]
        N = int(P.x_disc.shape[0])
        # algorithmic flops of the conditioning contraction: 2 * n_train * N per (candidate,
        # objective) (F = 2 n (N+1) + 2 n^2 + 2 n M; the contraction is the 2 n N term)
        flops_total = 3.0 * sum(2.0 * m.n * N * n_cand for m in P.model.models)
        achieved = flops_total / (ms_g * 1e-3) / 1e12
        dgemm_peak = measure_dgemm_peak(dev)
        tot_prof = sum(v[0] for v in prof.values())
        int8_engine = all(p.stats()[7] == 1 for p in plans) and not small_path
        if small_path:
            roofline = {
                "kernel": "small_kg_kernel (one CTA per candidate: kernel rows, K^-1 k_x, covariance row as fp64 dot "
                          "products, one warp per scalarisation marching over all lines)",
                "bound": "tensor", "achieved": achieved, "peak": dgemm_peak, "unit": "TFLOP/s",
                "frac": achieved / dgemm_peak,
                "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (burst, best of 5); at this size the path is "
                               "latency bound by construction (SURVEY 8d), the fraction is reported for completeness",
                "flops_per_launch": flops_total / max(n_g, 1), "avg_launch_ms": ms_g / max(n_g, 1),
                "share_of_step": ms_g / tot_prof if tot_prof > 0 else None, "traffic": None,
            }
        elif int8_engine:
            # The contraction runs on the int8 tensor cores as 28 exact digit-plane products
            # (7 balanced base-256 digits, 7 diagonals; csrc/dkg_ozaki.cu) of K padded to 32.  Its roofline is
            # the int8 tensor peak; B200's dense i8 rate is twice its bf16 rate, and the bf16 rate
            # is the measured cuBLAS figure of MEASURED_PEAKS.json (burst: this kernel is timed alone).
            bf16 = None
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    bf16 = float(json.load(f)["bf16_tflops"])
                src = "2 x MEASURED_PEAKS.json bf16_tflops (burst)"
            except Exception:
                bf16, src = 2250.0, "2 x 2250 TFLOP/s nominal dense bf16 (B200_PROFILING.md fallback)"
            # digit-pair products = pairs (i, j) of digits with i + j < diagonals: 7 balanced signed digits and 7
            # diagonals -> 28 (round 1: unsigned digits needed 8 diagonals -> 34); reduced precision 4 / 4 -> 10
            ns, ng = (7, int(os.environ.get("DKG_OZ_DIAGONALS", "7"))) if args.precision == "float64" else (4, 4)
            products = float(sum(1 for i in range(ns) for j in range(ns) if i + j < ng))
            kp = sum(-(-m.n // 32) * 32 for m in P.model.models) / sum(m.n for m in P.model.models)
            int8_ops = flops_total * products * kp
            int8_peak = 2.0 * bf16
            # the same kernel's MMA instruction stream with operand copies and accumulator drains switched
            # off: what the tensor pipe retires on THIS box under its power / clock conditions
            mma_peak = None
            if args.precision == "float64":
                try:
                    mma_peak, _ = _native.int8_peak(n_cand, N, max(m.n for m in P.model.models), 10)
                except Exception as exc:  # noqa: BLE001
                    extra["int8_peak_error"] = str(exc)
            peak = int8_peak / (products * kp)  # the int8 roofline in fp64-equivalent TFLOP/s
            roofline = {
                "kernel": "ozaki_kernel (tcgen05.mma.kind::i8 over base-256 digit planes: the fp64 GP "
                          f"conditioning contraction as {int(products)} exact int8 digit products)",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "peak_source": f"int8 tensor peak = {src} = {int8_peak:.0f} TOP/s, divided by the "
                               f"{products * kp:.1f} int8 ops the scheme spends per algorithmic fp64 flop",
                "int8_top_s_executed": int8_ops / (ms_g * 1e-3) / 1e12,
                "int8_peak_top_s": int8_peak,
                "int8_mma_peak_top_s_measured": mma_peak,
                "frac_vs_measured_int8_mma_peak": (int8_ops / (ms_g * 1e-3) / 1e12 / mma_peak) if mma_peak else None,
                "frac_vs_nominal_int8_4500_top_s": int8_ops / (ms_g * 1e-3) / 1e12 / 4500.0,
                "dgemm_peak_tflops": dgemm_peak,
                "frac_vs_dgemm_peak": achieved / dgemm_peak,
                "flops_per_launch": flops_total / max(n_g, 1),
                "avg_launch_ms": ms_g / max(n_g, 1),
                "share_of_step": ms_g / tot_prof if tot_prof > 0 else None,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel at this shape from the
                # ncu --set full capture profiles/r04c_obj0_kernels_ncu.csv (61.4 MB + 489.8 MB; algorithmic:
                # 537 MB of product rows written once, digit planes served from L2)
                "traffic": 551.3e6 if (n_cand == 4096 and N == 16384 and args.precision == "float64") else None,
                "traffic_source": "profiles/r04c_obj0_kernels_ncu.csv (ncu --set full, one launch, c4 shape)",
                "tensor_pipe_active_ncu_pct": 58.6 if (n_cand == 4096 and N == 16384 and args.precision == "float64") else None,
            }
        else:
            roofline = {
                "kernel": "dmma_gemm_kernel<cov> (GP conditioning contraction, fp64 DMMA)",
                "bound": "tensor", "achieved": achieved, "peak": dgemm_peak, "unit": "TFLOP/s",
                "frac": achieved / dgemm_peak,
                "peak_source": "cuBLAS DGEMM 8192^3 measured live in this run (burst, best of 5); "
                               "MEASURED_PEAKS.json has no fp64 figure",
                "flops_per_launch": flops_total / max(n_g, 1),
                "avg_launch_ms": ms_g / max(n_g, 1),
                "share_of_step": ms_g / tot_prof if tot_prof > 0 else None,
                "traffic": None,
            }
        peak = dgemm_peak
        extra["kernel_ms_per_step"] = {k: v[0] / 3.0 for k, v in prof.items()}
        extra["kernel_launches_per_step"] = {k: v[1] // 3 for k, v in prof.items()}
        whole = 1.03 * sum((2.0 * m.n * (N + 1) + 2.0 * m.n**2 + 2.0 * m.n * M) * n_cand for m in P.model.models)
        extra["path_roofline"] = {
            "algorithmic_flops_per_step_per_gpu": whole,
            "t_min_ms_at_peak": whole / (peak * 1e12) * 1e3,
            "frac_of_fp64_roofline": (whole / (peak * 1e12) * 1e3) / ms_step,
        }
        stats = [p.stats() for p in plans]
        extra["filter_stats"] = [
            {"survivors_per_set": s[1] / max(1, n_cand * S), "overflow_sets": s[2],
             "hull_vertices_per_set": s[3] / max(1, n_cand * S), "shortcut_sets": s[4]} for s in stats]

        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            om = _oracle_model(P)
            dense = _host_mem_gb() > 48.0 or N <= 4096
            t_acc, n_done, k = 0.0, 0, 0
            while t_acc < 12.0 and k < 64:
                t_acc += cpu_reference_step(P, om, k // 2, k % 2, dense)
                n_done += 1
                k += 1
            cpu_baseline = {
                "value": S * n_done / t_acc, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n_done} (candidate, objective) pairs x {S} scalarisations of the same workload, "
                          f"fwd+bwd, {'dense (N+1)^2 posterior' if dense else 'row-only posterior'} + Python hull "
                          f"loop per scalarisation ({t_acc:.1f} s of CPU work)",
            }
            # strong CPU baseline: row-only batched posterior (no dense covariance)
            t_acc2, n2 = 0.0, 0
            while t_acc2 < 6.0 and n2 < 64:
                t_acc2 += cpu_reference_step(P, om, n2 // 2, n2 % 2, False)
                n2 += 1
            extra["cpu_baseline_row_only"] = {"value": S * n2 / t_acc2, "unit": UNIT, "cores": cores,
                                              "kind": "port", "sample": f"{n2} pairs, row-only posterior"}
            # (V) of BASELINE.md section 3: vectorised, row-only, batched CPU implementation with the
            # candidate-independent work hoisted out (oracle/vectorised.py) -- the strong host baseline
            from oracle import vectorised as ov

            t0 = time.perf_counter()
            preps = [ov.Prepared(om, P.x_disc, P.weights, t) for t in range(M)]
            t_prep = time.perf_counter() - t0
            nb, t_acc3, n3 = 32, 0.0, 0
            while t_acc3 < 8.0 and n3 < 8:
                t0 = time.perf_counter()
                ov.kg_batch(preps[n3 % M], P.candidates[(n3 // M) * nb:(n3 // M + 1) * nb], need_grad=True)
                t_acc3 += time.perf_counter() - t0
                n3 += 1
            extra["cpu_baseline_vectorised"] = {
                "value": S * nb * n3 / t_acc3, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{n3} batches of {nb} candidates x 1 objective x {S} scalarisations, fwd+bwd, batched row-only "
                          f"posterior (one GEMM per batch), chord prefilter + exact march per set, envelope-theorem "
                          f"gradient; candidate-independent preparation ({t_prep:.2f} s for both objectives) not timed, "
                          f"as on the GPU",
            }

    if rank == 0:
        from decoupledbo_b200 import multi as _multi

        if _multi.TIMING is not None:  # DKG_MULTI_TIMING=1: diagnosis of the e2e path (numbers of such a run are not bench values)
            extra["multi_timing_ms"] = {k: 1e3 * v[0] / max(v[1], 1) for k, v in _multi.TIMING.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64" if args.precision == "float64" else "f64 (covariance contraction: 4 base-256 int8 digits)",
            "data": "synthetic",
            "config": workload_config(args.workload, P, per_gpu, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if strong is not None:
            line["strong"] = strong
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
