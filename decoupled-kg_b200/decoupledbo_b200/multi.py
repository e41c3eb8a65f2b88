"""Several acquisition functions at the same candidates in ONE pass.

The decoupled strategy evaluates one ``DiscreteKnowledgeGradient`` per objective
(``acquisition_optimisation_strategy.py:208``); the reference does so one after the other, one
candidate at a time.  Here the acquisition functions run concurrently -- one CUDA stream per plan,
so the latency-bound stages of one (hull march, backward gathers) overlap the streaming stages of
the other -- on one upload of the candidates, write into one packed result buffer, and -- under
``torch.distributed`` with a shard group -- each rank evaluates its contiguous share of the rows and
ONE all-gather (on the device, before the single device-to-host copy) gives every rank all values
and gradients of all objectives.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import distributed as _dist


class _Scratch:
    """Streams and staging buffers reused across calls (keyed by device / shapes)."""

    def __init__(self):
        self.streams = {}
        self.pinned = {}

    def streams_for(self, dev, n):
        key = (dev.index, n)
        if key not in self.streams:
            self.streams[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
        return self.streams[key]

    def pinned_next(self, shape):
        """Two pinned staging buffers per shape, handed out alternately: a result stays valid until the
        second-next call with the same shapes (a fresh pinned allocation costs ~0.2 ms per call: measured)."""
        pair = self.pinned.setdefault(tuple(shape), [None, None, 0])
        k = pair[2]
        if pair[k] is None:
            pair[k] = torch.empty(shape, dtype=torch.double).pin_memory()
        pair[2] = 1 - k
        return pair[k]


_scratch = _Scratch()

# DKG_MULTI_TIMING=1: per-phase host wall times (a device sync after every phase; diagnosis only)
import os as _os
import time as _time

TIMING = {} if _os.environ.get("DKG_MULTI_TIMING") == "1" else None


def _mark(name, t0, dev):
    if TIMING is None:
        return t0
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t1 = _time.perf_counter()
    acc = TIMING.setdefault(name, [0.0, 0])
    acc[0] += t1 - t0
    acc[1] += 1
    return t1


def evaluate_objectives(
    acqfs: Sequence, X: Tensor, need_grad: bool = False, group=None
) -> Tuple[Tensor, Optional[Tensor]]:
    """``acqfs``: M ``DiscreteKnowledgeGradient`` instances over the same input space;
    ``X``: ``(C, d)`` candidates (host or CUDA tensor; identical on every rank of ``group``).

    Returns ``kg (M, C)`` and, if ``need_grad``, ``dX (M, C, d)`` with ``dX[m, c] = d kg[m, c] / d X[c]``,
    on ``X``'s device, identical on every rank.  Both are views into one result block (each ``kg[m]`` and
    ``dX[m]`` is contiguous); for host inputs that block is one of two alternating pinned staging buffers, i.e. the
    results stay valid until the second-next call with the same shapes -- copy them to keep them longer."""
    plans = [a._get_plan() for a in acqfs]
    M = len(plans)
    dev = plans[0].device
    d = plans[0].d
    if X.dim() != 2 or X.shape[1] != d:
        raise ValueError(f"X must be (C, {d}); got {tuple(X.shape)}")
    on_host = not X.is_cuda
    C = X.shape[0]
    shard = group is not None and _dist.should_shard(C, group)
    world = dist.get_world_size(group) if shard else 1
    rank = dist.get_rank(group) if shard else 0
    lo, hi = _dist.shard_bounds(C, world, rank)
    rows = _dist.max_shard_rows(C, world)
    width = 1 + (d if need_grad else 0)
    import contextlib

    cuda = dev.type == "cuda"  # (CPU "plans" exist only in the gloo tests of this host logic)
    t0 = _time.perf_counter() if TIMING is not None else 0.0
    with (torch.cuda.device(dev) if cuda else contextlib.nullcontext()):
        Xs = X[lo:hi]
        if Xs.dtype != torch.double:
            Xs = Xs.to(torch.double)
        Xd = Xs.to(device=dev, non_blocking=True).contiguous()
        # packed per objective as [kg (rows) | dX (rows x d)] so that both are contiguous views
        buf = torch.empty(M, rows * width, dtype=torch.double, device=dev)
        if hi - lo < rows:
            buf.zero_()
        cur = torch.cuda.current_stream() if cuda else None
        streams = _scratch.streams_for(dev, M) if (cuda and M > 1) else [cur] * M
        n = hi - lo
        for m, (plan, s) in enumerate(zip(plans, streams)):
            if s is not cur:
                s.wait_stream(cur)
            with (torch.cuda.stream(s) if cuda else contextlib.nullcontext()):
                plan.forward_device(Xd, need_grad, out_kg=buf[m, :n],
                                    out_dX=buf[m, rows: rows + n * d] if need_grad else None)
        for s in streams:
            if s is not cur:
                cur.wait_stream(s)
        t0 = _mark("h2d+kernels", t0, dev)
        if shard:
            backend_cpu = dist.get_backend(group) == "gloo"
            send = buf.cpu() if backend_cpu else buf
            recv = torch.empty(world * send.numel(), dtype=torch.double, device=send.device)
            dist.all_gather_into_tensor(recv, send.reshape(-1), group=group)
            recv = recv.view(world, M, rows * width)
            # (world, M, [kg | dX]) -> one contiguous [kg (M, C) | dX (M, C, d)] block in row order, assembled
            # where the data is (on the device under NCCL): two strided copies, then ONE device-to-host copy
            res = torch.empty(M * C * width, dtype=torch.double, device=recv.device)
            kg_v = res[: M * C].view(M, C)
            dX_v = res[M * C:].view(M, C, d) if need_grad else None
            if C == world * rows:
                kg_v.view(M, world, rows).copy_(recv[:, :, :rows].permute(1, 0, 2))
                if need_grad:
                    dX_v.view(M, world, rows, d).copy_(recv[:, :, rows:].unflatten(2, (rows, d)).permute(1, 0, 2, 3))
            else:  # ragged shards
                for r in range(world):
                    rlo, rhi = _dist.shard_bounds(C, world, r)
                    k = rhi - rlo
                    kg_v[:, rlo:rhi] = recv[r, :, :k]
                    if need_grad:
                        dX_v[:, rlo:rhi] = recv[r, :, rows: rows + k * d].reshape(M, k, d)
            t0 = _mark("all_gather+assemble", t0, dev)
        else:
            res = buf.view(-1)  # rows == C: already [kg (C) | dX (C, d)] per objective
        if on_host and res.is_cuda:
            host = _scratch.pinned_next(tuple(res.shape))
            host.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            res = host
            t0 = _mark("d2h+sync", t0, dev)
        elif not on_host and not res.is_cuda:
            res = res.to(dev)
    if shard:
        kg = res[: M * C].view(M, C)
        dX = res[M * C:].view(M, C, d) if need_grad else None
    else:
        res = res.view(M, C * width)
        kg = res[:, :C]
        dX = res[:, C:].view(M, C, d) if need_grad else None
    return kg, dX
