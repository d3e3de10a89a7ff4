"""BASELINE config 4: VirConv-L dense-virtual stress — 50k..300k virtual points per scene, voxel cap lifted,
StVD input discard on vs off; rulebook-build throughput per layer type in input voxels/s and pairs/s, plus the time of
the GPU pre-backbone steps (StVD discard, voxelise + VFE).  Run on the GPU box:
    python profiles/rulebook_sweep.py > profiles/rulebook_sweep_r1.txt
CUDA-event timing, median of 7 repetitions, L2 flushed (256 MiB write) between repetitions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from virconv_b200 import ops, preprocess, scenes

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
BATCH = 2
SHAPE = [81, 1600, 1408]


def timed(fn, reps=7):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 1e3)
    return out, sorted(ts)[len(ts) // 2]


def stage_rulebooks(idx, shape, stride, proj, tag, rows):
    """rulebooks of one NRConv stage on index set `idx`: 3-D submanifold, projection + 2-D submanifold."""
    rb, t = timed(lambda: ops.build_subm_rulebook(idx, BATCH, shape, 3))
    rows.append((tag + ' subm3d', idx.shape[0], rb.n_pairs(), t))
    uv, t_uv = timed(lambda: ops.index2uv(idx, BATCH, proj, stride))
    rb2, t2 = timed(lambda: ops.build_subm_rulebook(uv, BATCH, [1600, 600], 3))
    rows.append((tag + ' index2uv', idx.shape[0], 0, t_uv))
    rows.append((tag + ' subm2d', idx.shape[0], rb2.n_pairs(), t2))


def run(n_virtual, discard):
    pts_list, calibs = [], []
    t_disc = 0.0
    for b, sid in enumerate((0, 1)):
        p = scenes.mask_points_by_range(scenes.make_points(sid, 16384, n_virtual)).astype(np.float32)
        t = torch.from_numpy(p).to(dev)
        if discard:
            np.random.seed(sid)
            (res, td) = timed(lambda: preprocess.split_and_discard(t, True, 0.8, False, np.random.RandomState(sid)), reps=3)
            t = res['points']
            t_disc += td
        pts_list.append(torch.cat([torch.full((t.shape[0], 1), float(b), device=dev), t], 1))
        calibs.append(scenes.Calib())
    pts = torch.cat(pts_list).contiguous()
    (feats, coords, num), t_vox = timed(lambda: ops.voxelize_mean(pts, BATCH, max_voxels=1_000_000))
    proj = ops.projection_params(calibs, None, BATCH, dev)
    rows = []
    idx, shape = coords.contiguous(), SHAPE
    stage_rulebooks(idx, shape, 1, proj, 'stage1', rows)
    for si, (stride, pad) in enumerate(((2, 1), (4, 1), (8, (0, 1, 1))), start=2):
        rb, t = timed(lambda: ops.build_conv_rulebook(idx, BATCH, shape, 3, 2, pad))
        rows.append((f'stage{si} conv3d s2', idx.shape[0], rb.n_pairs(), t))
        idx, shape = rb.out_indices, rb.out_shape
        stage_rulebooks(idx, shape, stride, proj, f'stage{si}', rows)
    rb, t = timed(lambda: ops.build_conv_rulebook(idx, BATCH, shape, (3, 1, 1), (2, 1, 1), 0))
    rows.append(('conv_out (3,1,1)', idx.shape[0], rb.n_pairs(), t))
    print(f'\n## {n_virtual} virtual points/scene, StVD input discard {"ON (rate 0.8, bin_num 2)" if discard else "off"}: '
          f'{pts.shape[0]} points in, {coords.shape[0]} voxels; discard {t_disc:.0f} us, voxelise+VFE {t_vox:.0f} us '
          f'({pts.shape[0] / t_vox:.1f} Mpoints/s)')
    tot_t = 0.0
    agg = {}
    for name, n, p, t in rows:
        print(f'  {name:18s} N_in={n:7d} pairs={p:9d}  {t:7.1f} us  {n / t:7.1f} Mvoxels/s' + (f'  {p / t:8.1f} Mpairs/s' if p else ''))
        tot_t += t
        kind = name.split()[-1] if 'conv' not in name else ('conv3d' if 'conv3d' in name or 'conv_out' in name else name.split()[-1])
        a = agg.setdefault(kind, [0, 0, 0.0])
        a[0] += n
        a[1] += p
        a[2] += t
    print(f'  all index work of one forward: {tot_t:.0f} us; per layer type: ' + '; '.join(
        f'{k}: {v[0] / v[2]:.0f} Mvoxels/s' + (f', {v[1] / v[2]:.0f} Mpairs/s' if v[1] else '') for k, v in agg.items()))


print('# rulebook-build throughput sweep (BASELINE config 4), batch of 2 scenes, grid [81,1600,1408], B200')
for nv in (50_000, 100_000, 200_000, 300_000):
    for disc in (False, True):
        run(nv, disc)
