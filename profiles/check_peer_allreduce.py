"""torchrun --nproc-per-node N profiles/check_peer_allreduce.py : the peer-memory all-reduce (csrc/allreduce.cu) against NCCL
on random buffers of several sizes (incl. sizes that are not multiples of 4 and the backbone's own 1.7 MB), repeated calls
(epoch parity / double buffering), and its time next to NCCL's."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from virconv_b200 import ops, parallel

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
dist.init_process_group('nccl', device_id=dev)
ok = True
for n in (7, 1024, 434177, 434176, 3000001):
    peer = parallel.PeerAllReduce(n, dev)
    for rep in range(5):
        g = torch.Generator(device=dev)
        g.manual_seed(1000 * rank + rep + n)
        x = torch.randn(n, device=dev, generator=g)
        ref = x.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        y = x.clone()
        peer(y, average=True)
        torch.cuda.synchronize()
        err = float((y - ref).abs().max())
        if err > 1e-5:
            ok = False
        if rank == 0:
            print(f'n={n} rep={rep} max|diff| vs NCCL {err:.2e}')
    # timing
    x = torch.randn(n, device=dev)
    for fn, name in ((lambda: peer(x, True), 'peer kernel'), (lambda: dist.all_reduce(x, op=dist.ReduceOp.AVG), 'NCCL AVG')):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 50], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f'n={n} ({4 * n / 1e6:.2f} MB) {name}: {float(t) * 1e3:.1f} us per call (max over {world} ranks, back to back)')
flag = int(ops.tc_error_flag(dev).item())
res = torch.tensor([1 if ok and flag == 0 else 0], device=dev)
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print('PEER ALLREDUCE', 'OK' if int(res) == 1 else 'FAILED', 'error flag', flag)
dist.destroy_process_group()
sys.exit(0 if int(res) == 1 else 1)
