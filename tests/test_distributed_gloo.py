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


# ---- evaluate_objectives (both objectives, one collective) and the SPMD optimiser ---------------
class _FakePlan:
    """Host stand-in for _native.Plan (forward_device writes into the caller's packed buffer)."""

    def __init__(self, shift):
        self.device = torch.device("cpu")
        self.d = 3
        self.shift = shift

    def forward_device(self, X, need_grad, out_kg=None, out_dX=None):
        kg, dX = _fake_eval(X + self.shift, need_grad)
        out_kg.copy_(kg)
        if need_grad:
            out_dX.copy_(dX.reshape(-1))
        return out_kg, out_dX


class _FakeAcqf:
    def __init__(self, shift):
        self._p = _FakePlan(shift)

    def _get_plan(self):
        return self._p


def _worker_multi(rank, world, port, C, ret):
    import numpy as np
    from scipy.optimize import minimize

    from decoupledbo_b200.multi import evaluate_objectives

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["DKG_SHARD_MIN_ROWS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        X = torch.rand(C, 3, generator=g, dtype=torch.double)
        acqfs = [_FakeAcqf(0.0), _FakeAcqf(0.25)]
        kg, dX = evaluate_objectives(acqfs, X, need_grad=True, group=dist.group.WORLD)
        ok = kg.shape == (2, C) and dX.shape == (2, C, 3)
        for m, a in enumerate(acqfs):
            wk, wd = _fake_eval(X + a._p.shift, True)
            ok = ok and torch.equal(kg[m], wk) and torch.equal(dX[m], wd)
        kg2, none = evaluate_objectives(acqfs, X, need_grad=False, group=dist.group.WORLD)
        ok = ok and none is None and torch.equal(kg2, kg)

        # SPMD L-BFGS: every rank runs the same scipy optimiser on the gathered (f, g)
        def f_and_g(x_np):
            Xc = torch.from_numpy(x_np.reshape(C, 3))
            v, gr = evaluate_objectives(acqfs[:1], Xc, need_grad=True, group=dist.group.WORLD)
            return float(v.sum()), gr[0].reshape(-1).numpy().copy()

        res = minimize(f_and_g, X.reshape(-1).numpy().copy(), jac=True, method="L-BFGS-B",
                       bounds=[(0.0, 1.0)] * (3 * C), options={"maxiter": 15})
        ret[rank] = (ok, res.x.tobytes(), res.nit)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("C", [9, 16])
def test_evaluate_objectives_and_spmd_lbfgs_world2(C):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_multi, args=(2, port, C, ret), nprocs=2, join=True)
    assert len(ret) == 2 and ret[0][0] and ret[1][0]
    assert ret[0][1] == ret[1][1] and ret[0][2] == ret[1][2] > 0  # identical iterates, no broadcast needed


def test_should_shard_thresholds(monkeypatch):
    from decoupledbo_b200 import distributed as D

    assert not D.should_shard(10**6)  # no process group
    monkeypatch.setenv("DKG_SHARD_MIN_ROWS", "100")
    assert D.min_rows_per_rank() == 100
