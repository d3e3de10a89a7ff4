"""Host-side cost of one bench step (cProfile over 10 steps); run on the GPU box."""
import cProfile
import pstats
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from virconv_b200 import scenes
from virconv_b200.backbone import VirConvL8x

dev = torch.device('cuda:0')
torch.manual_seed(666)
model = VirConvL8x(bench.CFG, 8, [1408, 1600, 80], precision='bf16').to(dev).train()
b = scenes.make_batch([0, 1], training=True)
vf, vc = torch.from_numpy(b.voxel_features).to(dev), torch.from_numpy(b.voxel_coords).to(dev)


def step():
    for p in model.parameters():
        p.grad = None
    out = model({'voxel_features': vf, 'voxel_coords': vc, 'batch_size': 2, 'calib': b.calib, 'aug_param': b.aug_param})
    loss = out['encoded_spconv_tensor'].features.mean()
    for t in out['multi_scale_3d_features'].values():
        loss = loss + t.features.mean()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
t = time.time()
for _ in range(10):
    step()
t_enq = time.time() - t
torch.cuda.synchronize()
t_all = time.time() - t
print(f'10 steps: enqueue-side wall {t_enq * 100:.2f} ms/step, with final sync {t_all * 100:.2f} ms/step')
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(22)
