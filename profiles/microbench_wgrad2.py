"""A/B of the tensor-core weight-gradient kernels on the bench workload, per layer (round 2).
    python profiles/microbench_wgrad2.py
variant 0 = round-1 kernel (128 persistent CTAs, one [K,Cin,Cout] partial each + reduce kernel), variant 1 = persistent
kernel of wgrad_tc2.cu (one scratch image, vector reductions, finalize).  Median of 7 CUDA-event timings per layer with the
L2 flushed in between; maximum relative difference between the variants (C = 8 layers: variant 1 against the fp32 kernel fed
bf16-rounded operands)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from virconv_b200 import _lib, executor, ops, scenes
from virconv_b200.backbone import VirConvL8x

dev = torch.device('cuda:0')
lib = _lib.load()
torch.manual_seed(666)
model = VirConvL8x(bench.CFG, 8, [1408, 1600, 80], precision='bf16').to(dev).train()
executor.ENABLED = False
b = scenes.make_batch([0, 1], training=True)
calls = []
orig = ops.conv_forward


def spy(feats, weight, rb, bn_sums=None, precision='fp32', feats_bf16=None, keep=None):
    calls.append((feats.detach(), weight.detach(), rb))
    return orig(feats, weight, rb, bn_sums, precision, feats_bf16, keep)


ops.conv_forward = spy
with torch.no_grad():
    model({'voxel_features': torch.from_numpy(b.voxel_features).to(dev), 'voxel_coords': torch.from_numpy(b.voxel_coords).to(dev),
           'batch_size': 2, 'calib': b.calib, 'aug_param': b.aug_param})
ops.conv_forward = orig
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_p, _ws, _stream = ops._p, ops._ws, ops._stream
err = ops.tc_error_flag(dev)


def timed(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    return sorted(ts)[len(ts) // 2]


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float32)


tot = {}
for li, (feats, weight, rb) in enumerate(calls):
    cout, cin = weight.shape[0], weight.shape[-1]
    feats = feats.contiguous()
    fb = ops.cast_bf16(feats)
    dout = torch.randn(rb.n_out, cout, device=dev)
    db = ops.cast_bf16(dout)
    wsb = lib.vc_conv_wgrad_tc_ws_bytes(rb.n_out, cin, cout, rb.K)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    old_ok = ops.tc_supported(cin, cout)
    res = {}

    def wgrad(variant):
        dw = torch.empty(tuple(weight.shape), dtype=torch.float32, device=dev)

        def run():
            _lib.check(lib.vc_conv_wgrad_tc(_p(fb), _p(db), _p(rb.nbr), _p(dw), rb.n_out, cin, cout, rb.K, _p(ws), ws.numel(),
                                            _p(err), _stream()), 'vc_conv_wgrad_tc')
        _lib.check(lib.vc_set_tc_variant(variant), 'variant')
        t = timed(run)
        return t, dw

    res['wgrad1'], w1 = wgrad(1)
    if old_ok:
        res['wgrad0'], w0 = wgrad(0)
    else:
        lib.vc_set_tc_variant(1)
        w0 = ops.conv_wgrad(bf16r(feats), bf16r(dout), weight.shape, rb, 'fp32')
    lib.vc_set_tc_variant(1)
    d = rel(w1, w0)
    for k, v in res.items():
        tot[k] = tot.get(k, 0.0) + v
    P = rb.n_pairs()
    by = (feats.shape[0] * cin + rb.n_out * cout) * 2 + P * 8 + rb.K * cin * cout * 4
    print(f'L{li:02d} {"subm" if rb.subm else "conv"}{rb.ndim}d {cin:2d}->{cout:2d} K={rb.K:2d} N_out={rb.n_out:6d} P={P:8d} '
          + ' '.join(f'{k}={v:6.1f}us' for k, v in sorted(res.items()))
          + f'  wgrad1: {by / res["wgrad1"] / 1e3:7.1f} GB/s {2 * P * cin * cout / res["wgrad1"] / 1e6:6.1f} TF/s  maxrel {d:.1e}', flush=True)
print('TOTAL us', {k: round(v, 1) for k, v in sorted(tot.items())})
print('error flag', int(err.item()))
