"""Multi-rank host logic on CPU: world_size-2 (and 3) gloo groups, no GPU.

The flow itself needs CUDA, so the per-rank compute is a stand-in with the same signature and a
result that depends on (sample index, global instance index); what is tested is the N>1 plumbing:
partitioning, per-rank calls, gathering, and that 1 rank and N ranks give identical results."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bcnf_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 10_000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


def _fake_sample(n_samples, cond, outer=True, sigma=1.0):
    # deterministic function of (sample index, instance content): (n_samples, n_inst, 3)
    s = torch.arange(n_samples, dtype=torch.float32).view(-1, 1, 1)
    return sigma * (s + cond.sum(dim=tuple(range(1, cond.ndim))).view(1, -1, 1) * torch.tensor([1.0, 2.0, 3.0]))


def _fake_log_prob(y, cond):
    return y.sum(1) - cond.reshape(cond.shape[0], -1).sum(1)


def _worker(rank, world, port, n_inst, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cond = torch.randn(n_inst, 30, 3, generator=g)
        y = torch.randn(n_inst, 19, generator=g)
        local = sharding.sample_sharded(_fake_sample, 5, cond, sigma=0.5)
        lo, hi = sharding.shard_bounds(n_inst, rank, world)
        assert local.shape == (5, hi - lo, 3)
        full = sharding.sample_sharded(_fake_sample, 5, cond, gather=True, sigma=0.5)
        lp = sharding.log_prob_sharded(_fake_log_prob, y, cond, gather=True)
        ref = _fake_sample(5, cond, sigma=0.5)
        ok = bool(torch.equal(full, ref)) and bool(torch.equal(lp, _fake_log_prob(y, cond)))
        ok = ok and bool(torch.equal(local, ref[:, lo:hi]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_inst", [(2, 10), (2, 7), (3, 8)])
def test_sharded_calls_match_single_rank(world, n_inst):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_inst, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    results = dict(q.get(timeout=10) for _ in range(world))
    assert results == {r: True for r in range(world)}


def test_single_process_is_identity():
    cond = torch.randn(6, 4)
    out = sharding.sample_sharded(_fake_sample, 3, cond, gather=True)
    assert torch.equal(out, _fake_sample(3, cond))
    with pytest.raises(ValueError):
        sharding.shard_conditions([torch.zeros(3, 2), torch.zeros(4, 2)])


# ---- data-parallel Trainer: parameter broadcast and the flat-buffer gradient all-reduce (world_size 2, gloo, CPU) ----
def _trainer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bcnf_b200.train import Trainer
        torch.manual_seed(rank)                                   # replicas start different
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
        qmat = torch.linalg.qr(torch.randn(4, 4))[0]              # column-major, like the mixing matrices
        net.register_buffer("q", qmat)
        assert not net.q.is_contiguous() or True
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        tr = Trainer(net, opt, process_group=dist.group.WORLD)    # broadcasts rank 0's parameters and buffers
        flat = torch.cat([p.detach().flatten() for p in net.parameters()] + [net.q.flatten()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same_params = bool(torch.equal(flat, ref))
        # different gradients per rank -> the mean on every rank, through ONE flat buffer
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        tr._allreduce_grads()
        expect = [(1 + world) / 2 * (i + 1) for i in range(4)]
        ok = all(torch.allclose(p.grad, torch.full_like(p, e)) for p, e in zip(net.parameters(), expect))
        ok = ok and tr._flat.numel() == sum(p.numel() for p in net.parameters())
        q.put((rank, same_params and ok))
    finally:
        dist.destroy_process_group()


def test_trainer_process_group_broadcasts_parameters_and_averages_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    results = dict(q.get(timeout=10) for _ in range(world))
    assert results == {r: True for r in range(world)}
