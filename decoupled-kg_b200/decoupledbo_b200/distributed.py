"""Multi-GPU sharding of the hot path.

Candidates (t-batch rows / optimiser restarts) are independent (reference ``discretekg.py:145``),
so they shard contiguously across ranks with replicated GP state; the only collective is one
all-gather of the acquisition values (and gradients when requested) so that every rank can take
the same first-index argmax.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests).  Nothing here touches the CUDA library directly: the per-shard evaluation is a
callable, so the host logic is testable on CPU.
"""

from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced shards: the first ``n_rows % world_size`` ranks get one extra row."""
    base, extra = divmod(n_rows, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def max_shard_rows(n_rows: int, world_size: int) -> int:
    return -(-n_rows // world_size)


def is_active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_group():
    return dist.group.WORLD


def min_rows_per_rank() -> int:
    """A t-batch is sharded only when every rank gets at least this many rows (DKG_SHARD_MIN_ROWS,
    default 64): below that the forward is launch-bound and every rank simply evaluates all rows
    (identical results on every rank, no collective)."""
    return max(1, int(os.environ.get("DKG_SHARD_MIN_ROWS", "64")))


def should_shard(n_rows: int, group=None) -> bool:
    if not (dist.is_available() and dist.is_initialized()):
        return False
    world = dist.get_world_size(group)
    return world > 1 and n_rows >= world * min_rows_per_rank()


def _gather_rows(send: Tensor, n_rows: int, world: int, group) -> Tensor:
    """all-gather of per-rank row blocks ``send`` (rows_max, width) -> (n_rows, width) in row order.
    NCCL gathers on the device; gloo (CPU tests) needs host tensors."""
    on_cpu_backend = dist.get_backend(group) == "gloo"
    buf = send.cpu() if (on_cpu_backend and send.is_cuda) else send
    rows = buf.shape[0]
    recv = torch.empty(world * rows, buf.shape[1], dtype=buf.dtype, device=buf.device)
    dist.all_gather_into_tensor(recv, buf.contiguous(), group=group)
    base, extra = divmod(n_rows, world)
    if extra == 0:
        full = recv  # every shard is full: already in row order
    else:
        recv = recv.view(world, rows, -1)
        full = torch.cat([recv[r, : base + (1 if r < extra else 0)] for r in range(world)], dim=0)
    return full.to(send.device) if full.device != send.device else full


def sharded_evaluate(
    X: Tensor,
    evaluate: Callable[[Tensor, bool], Tuple[Tensor, Optional[Tensor]]],
    need_grad: bool = False,
    group=None,
) -> Tuple[Tensor, Optional[Tensor]]:
    """Evaluate ``evaluate(X_shard, need_grad) -> (kg, dX)`` on this rank's rows of ``X`` (C, d)
    and all-gather: every rank returns the full ``kg`` (C,) and ``dX`` (C, d) in row order.
    ONE collective per call: values and gradients travel in the same (rows, 1 + d) block."""
    if not (dist.is_available() and dist.is_initialized()):
        return evaluate(X, need_grad)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    C, d = X.shape
    lo, hi = shard_bounds(C, world, rank)
    kg_loc, dX_loc = evaluate(X[lo:hi], need_grad)
    rows = max_shard_rows(C, world)
    width = 1 + (d if need_grad else 0)
    send = torch.zeros(rows, width, dtype=kg_loc.dtype, device=kg_loc.device)
    send[: hi - lo, 0] = kg_loc
    if need_grad:
        send[: hi - lo, 1:] = dX_loc
    full = _gather_rows(send, C, world, group)
    kg = full[:, 0].contiguous()
    dX = full[:, 1:].contiguous() if need_grad else None
    return kg, dX


def sharded_forward(plan, X: Tensor, need_grad: bool, group=None) -> Tuple[Tensor, Optional[Tensor]]:
    """The product path behind ``DiscreteKnowledgeGradient.forward`` when a shard group is set: this
    rank's contiguous rows of ``X`` go through ``plan`` on the GPU, the (kg | dX) blocks are
    all-gathered ON THE DEVICE (NCCL), and only then -- if ``X`` was a host tensor -- copied back
    once.  Every rank must call this with the same ``X``."""
    on_host = not X.is_cuda

    def evaluate(Xs: Tensor, ng: bool):
        Xd = Xs.to(device=plan.device, dtype=torch.double, non_blocking=True) if on_host else Xs
        return plan.forward_device(Xd.contiguous(), ng)

    kg, dX = sharded_evaluate(X, evaluate, need_grad, group)
    if on_host:
        kg = kg.cpu()
        dX = dX.cpu() if dX is not None else None
    return kg, dX


def first_argmax(values: Tensor) -> int:
    """First index of the maximum (what ``torch.argmax`` / BoTorch's restart selection use);
    identical on every rank because the gathered values are identical."""
    return int(torch.argmax(values))
