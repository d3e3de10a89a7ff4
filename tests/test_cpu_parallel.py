"""CPU tests of the N>1 path (gloo, world_size 2): scene sharding and the single-bucket gradient all-reduce that
bench.py runs under torchrun with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from virconv_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical weights on every rank (DDP broadcasts them)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 4))
    torch.manual_seed(100 + rank)                           # different data per rank (different scenes)
    x = torch.randn(32, 8)
    model(x).square().mean().backward()
    local = [p.grad.clone() for p in model.parameters()]
    nbytes = parallel.allreduce_gradients(list(model.parameters()), average=True)
    q.put((rank, [g.numpy() for g in local], [p.grad.clone().numpy() for p in model.parameters()], nbytes))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, r0, b0), (_, l1, r1, b1) = res
    n_params = sum(g.size for g in l0)
    assert b0 == b1 == 4 * n_params                         # one bucket holding every gradient
    for a, b, m0, m1 in zip(l0, l1, r0, r1):
        want = (torch.from_numpy(a) + torch.from_numpy(b)) / 2
        assert torch.allclose(torch.from_numpy(m0), want, atol=1e-7)
        assert torch.equal(torch.from_numpy(m0), torch.from_numpy(m1))   # every rank ends with the same gradient


def test_scene_sharding_is_a_partition():
    world, per = 4, 2
    for step in range(3):
        ids = [parallel.shard_scene_ids(step, r, world, per) for r in range(world)]
        flat = sorted(i for s in ids for i in s)
        assert flat == list(range(step * world * per, (step + 1) * world * per))
        assert all(len(s) == per for s in ids)


def test_allreduce_is_noop_without_process_group():
    m = torch.nn.Linear(4, 4)
    m(torch.randn(2, 4)).sum().backward()
    g = m.weight.grad.clone()
    assert parallel.allreduce_gradients(list(m.parameters())) == 0
    assert torch.equal(g, m.weight.grad)
