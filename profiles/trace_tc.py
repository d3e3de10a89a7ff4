"""Pipeline timeline of the tcgen05 gather-GEMM (debug build -DVC_TC_TRACE, VIRCONV_LIB=..._trace.so): per CTA SM-clock
timestamps: start, mainloop start, every stage's `full` wake-up in the MMA warp, accumulator done, CTA end."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from virconv_b200 import _lib, ops, scenes
from virconv_b200.backbone import VirConvL8x

dev = torch.device('cuda:0')
torch.manual_seed(666)
model = VirConvL8x(bench.CFG, 8, [1408, 1600, 80], precision='bf16').to(dev).train()
b = scenes.make_batch([0, 1], training=True)
calls = []
orig = ops.conv_forward


def spy(feats, weight, rb, bn_sums=None, precision='fp32', feats_bf16=None, keep=None):
    calls.append((feats.detach(), weight.detach(), rb, precision))
    return orig(feats, weight, rb, bn_sums, precision, feats_bf16, keep)


ops.conv_forward = spy
with torch.no_grad():
    model({'voxel_features': torch.from_numpy(b.voxel_features).to(dev), 'voxel_coords': torch.from_numpy(b.voxel_coords).to(dev),
           'batch_size': 2, 'calib': b.calib, 'aug_param': b.aug_param})
ops.conv_forward = orig
lib = _lib.load()
lib.vc_debug_set_trace.argtypes = [ctypes.c_void_p]
for li in (6, 10, 11):
    feats, weight, rb, prec = calls[li]
    fb = ops.cast_bf16(feats)
    ops.conv_forward(feats, weight, rb, None, prec, fb)
    trace = torch.zeros(64 * 64, dtype=torch.int64, device=dev)
    assert lib.vc_debug_set_trace(trace.data_ptr()) == 0
    ops.conv_forward(feats, weight, rb, None, prec, fb)
    torch.cuda.synchronize()
    lib.vc_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(64, 64)
    print(f'layer {li}: {weight.shape[-1]}->{weight.shape[0]} N={rb.n_out} (SM clock cycles relative to CTA start; ~1.9 GHz)')
    for cta in (0, 1, 17, 40):
        r = t[cta]
        nk = int(r[62])
        st = r[4:4 + min(nk, 27)] - r[0]
        print(f'  cta {cta}: nk={nk} init={r[1]-r[0]} mainloop_start={r[2]-r[0]} stage wake-ups d={np.diff(st).tolist()} first={st[0] if len(st) else None} '
              f'accum_done={r[40]-r[0]} end={r[41]-r[0]}')
