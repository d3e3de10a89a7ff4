"""Where one bench step goes: host enqueue time vs GPU time vs GPU busy time (sum of kernel durations from CUPTI via
torch.profiler).  Run on the GPU box:  python profiles/step_anatomy.py [bf16|fp32]"""
import os
import sys
import time
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from virconv_b200 import scenes
from virconv_b200.backbone import VirConvL8x

prec = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
dev = torch.device('cuda:0')
torch.manual_seed(666)
model = VirConvL8x(bench.CFG, 8, [1408, 1600, 80], precision=prec).to(dev).train()
params = list(model.parameters())
b = scenes.make_batch([0, 1], bench.N_LIDAR, bench.N_VIRTUAL, bench.MAX_VOXELS, training=True)
vf, vc = torch.from_numpy(b.voxel_features).to(dev), torch.from_numpy(b.voxel_coords).to(dev)


def step():
    for p in params:
        p.grad = None
    out = model({'voxel_features': vf, 'voxel_coords': vc, 'batch_size': 2, 'calib': b.calib, 'aug_param': b.aug_param,
                 'virconv_inputs_ready': True})
    loss = out['encoded_spconv_tensor'].features.mean()
    for t in out['multi_scale_3d_features'].values():
        loss = loss + t.features.mean()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 20
t = time.perf_counter()
for _ in range(N):
    step()
t_enq = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f'{N} steps back to back: host enqueue {t_enq / N * 1e3:.2f} ms/step, wall incl. final sync {t_all / N * 1e3:.2f} ms/step')

# host enqueue time when the GPU is never the bottleneck of the python thread: sync before every step
hs = []
gs = []
for _ in range(N):
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    a.record()
    step()
    e.record()
    hs.append(time.perf_counter() - t)
    torch.cuda.synchronize()
    gs.append(a.elapsed_time(e))
print(f'isolated steps: host enqueue {1e3 * sum(hs) / N:.2f} ms, first-to-last GPU event {sum(gs) / N:.2f} ms')

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
busy = defaultdict(lambda: [0, 0.0])
tot = 0.0
first, last = None, None
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        d = ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
        busy[ev.name[:60]][0] += 1
        busy[ev.name[:60]][1] += d
        tot += d
print(f'GPU busy (sum of kernel+memcpy durations): {tot / 3 / 1e3:.3f} ms/step over {sum(v[0] for v in busy.values()) / 3:.0f} device ops')
for k, v in sorted(busy.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f'  {v[1] / 3:9.1f} us  n={v[0] / 3:5.1f}  {k}')

import cProfile
import pstats
from virconv_b200 import _lib
lib = _lib.load()
acc = {'fwd': 0.0, 'bwd': 0.0}
_f, _b = lib.vc_exec_forward, lib.vc_exec_backward


class _Wrap:
    def __init__(self, fn, key):
        self.fn, self.key = fn, key

    def __call__(self, *a):
        t0 = time.perf_counter()
        r = self.fn(*a)
        acc[self.key] += time.perf_counter() - t0
        return r


lib.vc_exec_forward = _Wrap(_f, 'fwd')
lib.vc_exec_backward = _Wrap(_b, 'bwd')
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(N):
    step()
t_enq = time.perf_counter() - t
torch.cuda.synchronize()
print(f'host time inside vc_exec_forward {acc["fwd"] / N * 1e3:.2f} ms/step, vc_exec_backward {acc["bwd"] / N * 1e3:.2f} ms/step, '
      f'whole step enqueue {t_enq / N * 1e3:.2f} ms')
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(25)

# per-stream timeline of one step (kineto): busy time per stream, span, and the gaps on the busiest stream
with profile(activities=[ProfilerActivity.CUDA]) as prof2:
    torch.cuda.synchronize()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = []
for e in prof2.profiler.kineto_results.events():
    if e.device_type() == torch.autograd.DeviceType.CUDA and e.duration_ns() > 0:
        evs.append((e.start_ns(), e.start_ns() + e.duration_ns(), e.device_resource_id(), e.name()))
evs.sort()
if evs:
    span = (evs[-1][1] - evs[0][0]) / 3e6
    per = defaultdict(float)
    for a, b_, sid, _ in evs:
        per[sid] += (b_ - a) / 3e6
    # union coverage
    cov, cur_a, cur_b = 0, None, None
    for a, b_, _, _ in evs:
        if cur_b is None or a > cur_b:
            if cur_b is not None:
                cov += cur_b - cur_a
            cur_a, cur_b = a, b_
        else:
            cur_b = max(cur_b, b_)
    cov += cur_b - cur_a
    print(f'timeline: span {span:.3f} ms/step, any-stream busy {cov / 3e6:.3f} ms/step, per stream busy ms/step: '
          + ', '.join(f'{k}: {v:.3f}' for k, v in sorted(per.items(), key=lambda kv: -kv[1])))
    main_id = max(per, key=per.get)
    me = [x for x in evs if x[2] == main_id]
    gaps = [(me[i + 1][0] - me[i][1]) / 1e3 for i in range(len(me) - 1)]
    gaps_small = [g for g in gaps if 0 < g < 50]
    print(f'busiest stream: {len(me) / 3:.0f} ops/step, gaps <50us: n={len(gaps_small) / 3:.0f}/step, sum {sum(gaps_small) / 3e3:.3f} ms/step, '
          f'median {sorted(gaps_small)[len(gaps_small) // 2] if gaps_small else 0:.1f} us')
