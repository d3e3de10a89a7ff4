"""Per-layer timing of the conv kernels on the bench workload (CUDA events, L2 flushed between repetitions).
    python profiles/microbench_conv.py [fwd|all]
Prints one line per conv layer: shape, pairs, us, achieved GB/s (algorithmic bytes) and TFLOP/s."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from virconv_b200 import ops, scenes, spconv_compat as sp
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
mode = sys.argv[1] if len(sys.argv) > 1 else 'fwd'
for li, (feats, weight, rb, prec) in enumerate(calls):
    cout, cin = weight.shape[0], weight.shape[-1]
    fb = ops.cast_bf16(feats) if ops.tc_supported(cin, cout) else None
    dout = torch.randn(rb.n_out, cout, device=dev)
    db = ops.cast_bf16(dout) if fb is not None else None
    res = {}
    fns = {'fwd': lambda: ops.conv_forward(feats, weight, rb, None, prec, fb)}
    if mode == 'all':
        fns['dgrad'] = lambda: ops.conv_dgrad(dout, weight, rb, prec, db)
        fns['wgrad'] = lambda: ops.conv_wgrad(feats, dout, weight.shape, rb, prec, fb, db)
    for name, fn in fns.items():
        fn()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) * 1e3)
        res[name] = sorted(ts)[len(ts) // 2]
        tot[name] += res[name]
    P = rb.n_pairs()
    by = feats.shape[0] * cin * 2 + rb.n_out * cout * 4 + P * 8 + rb.K * cin * cout * 2
    print(f'L{li:02d} {"subm" if rb.subm else "conv"}{rb.ndim}d {cin:2d}->{cout:2d} K={rb.K:2d} N_in={feats.shape[0]:6d} N_out={rb.n_out:6d} '
          f'P={P:8d} ' + ' '.join(f'{k}={v:7.1f}us' for k, v in res.items()) +
          f'  fwd: {by / res["fwd"] / 1e3:7.1f} GB/s {2 * P * cin * cout / res["fwd"] / 1e6:6.1f} TF/s')
print('TOTAL us', {k: round(v, 1) for k, v in tot.items()}, 'lib', os.environ.get('VIRCONV_LIB', 'default'))
