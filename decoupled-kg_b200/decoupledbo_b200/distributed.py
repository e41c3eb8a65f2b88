"""Multi-GPU sharding of the hot path.

Candidates (t-batch rows / optimiser restarts) are independent (reference ``discretekg.py:145``),
so they shard contiguously across ranks with replicated GP state; the only collective is one
all-gather of the acquisition values (and gradients when requested) so that every rank can take
the same first-index argmax.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in
the CPU tests).  Nothing here touches the CUDA library directly: the per-shard evaluation is a
callable, so the host logic is testable on CPU.
"""

from __future__ import annotations

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


def sharded_evaluate(
    X: Tensor,
    evaluate: Callable[[Tensor, bool], Tuple[Tensor, Optional[Tensor]]],
    need_grad: bool = False,
    group=None,
) -> Tuple[Tensor, Optional[Tensor]]:
    """Evaluate ``evaluate(X_shard, need_grad) -> (kg, dX)`` on this rank's rows of ``X`` (C, d)
    and all-gather: every rank returns the full ``kg`` (C,) and ``dX`` (C, d) in row order."""
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
    recv = torch.empty(world * rows, width, dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, rows, width)
    parts = []
    for r in range(world):
        rlo, rhi = shard_bounds(C, world, r)
        parts.append(recv[r, : rhi - rlo])
    full = torch.cat(parts, dim=0)
    kg = full[:, 0].contiguous()
    dX = full[:, 1:].contiguous() if need_grad else None
    return kg, dX


def first_argmax(values: Tensor) -> int:
    """First index of the maximum (what ``torch.argmax`` / BoTorch's restart selection use);
    identical on every rank because the gathered values are identical."""
    return int(torch.argmax(values))
