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


def _worker(rank, world, port, q, flat_grads=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical weights on every rank (DDP broadcasts them)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 4))
    torch.manual_seed(100 + rank)                           # different data per rank (different scenes)
    x = torch.randn(32, 8)
    model(x).square().mean().backward()
    if flat_grads:
        # what the plan executor hands autograd: every .grad is a slice of ONE flat buffer -> reduced in place
        params = list(model.parameters())
        buf = torch.cat([p.grad.reshape(-1) for p in params])
        for p, v in zip(params, buf.split([p.numel() for p in params])):
            p.grad = v.view_as(p)
        assert parallel._as_one_buffer([p.grad for p in params]) is not None
    local = [p.grad.clone() for p in model.parameters()]
    nbytes = parallel.allreduce_gradients(list(model.parameters()), average=True)
    if flat_grads:
        assert all(p.grad.untyped_storage().data_ptr() == buf.untyped_storage().data_ptr() for p in params)
    q.put((rank, [g.numpy() for g in local], [p.grad.clone().numpy() for p in model.parameters()], nbytes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('flat_grads', [False, True])
def test_gradient_allreduce_two_ranks_gloo(flat_grads):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, flat_grads)) for r in range(world)]
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
