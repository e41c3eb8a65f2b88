"""world_size-2 gloo test of the candidate sharding + all-gather host logic (CPU)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decoupledbo_b200.distributed import first_argmax, shard_bounds, sharded_evaluate


def test_shard_bounds_cover_everything():
    for C in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(C, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_eval(X, need_grad):
    kg = (X**2).sum(-1) + torch.sin(3 * X[:, 0])
    return kg, (2 * X if need_grad else None)


def _worker(rank, world, port, C, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        X = torch.rand(C, 3, generator=g, dtype=torch.double)
        kg, dX = sharded_evaluate(X, _fake_eval, need_grad=True)
        want_kg, want_dX = _fake_eval(X, True)
        ok = torch.equal(kg, want_kg) and torch.equal(dX, want_dX)
        kg2, none = sharded_evaluate(X, _fake_eval, need_grad=False)
        ok = ok and none is None and torch.equal(kg2, want_kg)
        ret[rank] = (ok, first_argmax(kg), first_argmax(want_kg))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("C", [9, 64])
def test_sharded_evaluate_world2(C):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, C, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for r in range(2):
        ok, got, want = ret[r]
        assert ok and got == want
    assert ret[0][1] == ret[1][1]  # every rank takes the same argmax
