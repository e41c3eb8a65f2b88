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
    on ``X``'s device, identical on every rank.  Both are VIEWS into one result block (not necessarily
    contiguous); for host inputs that block is one of two alternating pinned staging buffers, i.e. the
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
            # per-candidate records [objective][kg | dX]: a rank's shard is then ONE contiguous block of the
            # gathered buffer, which is already in row order -- no reordering after the collective
            pack = torch.empty(rows, M, width, dtype=torch.double, device=dev)
            if n < rows:
                pack.zero_()
            pack[:n, :, 0] = buf[:, :n].t()
            if need_grad:
                pack[:n, :, 1:] = buf[:, rows: rows + n * d].view(M, n, d).permute(1, 0, 2)
            backend_cpu = dist.get_backend(group) == "gloo"
            send = pack.cpu() if backend_cpu else pack
            recv = torch.empty(world * rows * M * width, dtype=torch.double, device=send.device)
            dist.all_gather_into_tensor(recv, send.reshape(-1), group=group)
            recv = recv.view(world * rows, M, width)
            if C != world * rows:  # ragged shards: drop the padding rows of the shorter ones
                recv = torch.cat([recv[r * rows: r * rows + (_dist.shard_bounds(C, world, r)[1] - _dist.shard_bounds(C, world, r)[0])]
                                  for r in range(world)], dim=0)
            t0 = _mark("pack+all_gather", t0, dev)
            res, packed = recv, True
        else:
            res, packed = buf, False
        if on_host and res.is_cuda:
            host = _scratch.pinned_next(tuple(res.shape))
            host.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            res = host
            t0 = _mark("d2h+sync", t0, dev)
        elif not on_host and not res.is_cuda:
            res = res.to(dev)
    # views into the one result block (possibly non-contiguous)
    if packed:  # (C, M, width)
        kg = res[:, :, 0].t()
        dX = res[:, :, 1:].permute(1, 0, 2) if need_grad else None
    else:       # (M, rows * width) with rows == C
        kg = res[:, :C]
        dX = res[:, C:].view(M, C, d) if need_grad else None
    return kg, dX
