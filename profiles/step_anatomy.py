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
