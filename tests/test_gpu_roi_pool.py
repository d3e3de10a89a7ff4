"""GPU parity tests of the voxel-RoI pooling primitives against (1) the REFERENCE's own CUDA
kernels, compiled from /root/reference into oracle/_ref by oracle/ref_build.py (bit-exact, realistic sizes), and (2) the
numpy restatement oracle/pointnet2.py (small cases, also covers empty balls / ragged batches)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import pointnet2 as o_pn
from oracle import ref_build

pytestmark = pytest.mark.gpu


def _scene(seed, n_per=(5000, 4200), shape=(21, 400, 352), n_query=(3000, 3000), stride=4, jitter=0.5):
    """Sparse voxels (batch-contiguous rows), their centres, the dense voxel->row map and query points near them."""
    rng = np.random.default_rng(seed)
    vs = np.array([0.05, 0.05, 0.05], np.float32) * stride
    lo = np.array([0, -40, -3], np.float32)
    coords, cnt = [], []
    for b, n in enumerate(n_per):
        # n distinct cells inside a window holding ~8n cells (density like a LiDAR surface patch), rows sorted by cell
        wy = int(min(shape[1], max(4, round((8 * n / shape[0]) ** 0.5))))
        wx = int(min(shape[2], max(4, -(-8 * n // (shape[0] * wy)))))
        n = min(n, shape[0] * wy * wx)
        win = np.sort(rng.choice(shape[0] * wy * wx, size=n, replace=False))
        z, y, x = win // (wy * wx), (win // wx) % wy + (shape[1] - wy) // 2, win % wx + (shape[2] - wx) // 3
        coords.append(np.stack([np.full_like(z, b), z, y, x], 1))
        cnt.append(n)
    coords = np.concatenate(coords).astype(np.int32)
    xyz = np.ascontiguousarray(((coords[:, [3, 2, 1]].astype(np.float32) + 0.5) * vs + lo).astype(np.float32))
    v2p = -np.ones((len(n_per),) + tuple(shape), dtype=np.int32)
    v2p[coords[:, 0], coords[:, 1], coords[:, 2], coords[:, 3]] = np.arange(len(coords), dtype=np.int32)
    new_xyz, new_coords = [], []
    start = 0
    for b, (n, m) in enumerate(zip(cnt, n_query)):
        pick = rng.integers(0, n, m) + start
        p = xyz[pick] + rng.normal(0, jitter, (m, 3)).astype(np.float32)
        c = np.ascontiguousarray(np.floor((p - lo) / vs).astype(np.int32)[:, [2, 1, 0]])
        new_xyz.append(p)
        new_coords.append(np.concatenate([np.full((m, 1), b, np.int32), c], 1))
        start += n
    return (xyz, np.array(cnt, np.int32), np.ascontiguousarray(np.concatenate(new_xyz).astype(np.float32)),
            np.array(n_query, np.int32), np.ascontiguousarray(np.concatenate(new_coords).astype(np.int32)), v2p)


def _ref():
    if not ref_build.available():
        pytest.skip('oracle/_ref/libpointnet2_stack_ref.so not built (needs /root/reference at build time)')
    lib = ctypes.CDLL(ref_build.LIB)
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize('max_range,radius,nsample', [((4, 4, 4), 0.8, 16), ((2, 2, 2), 0.4, 16), ((1, 3, 5), 1.6, 5)])
def test_voxel_query_and_grouping_vs_compiled_reference(lib_built, max_range, radius, nsample):
    from virconv_b200 import roi_pool
    lib = _ref()
    xyz, xyz_cnt, new_xyz, new_cnt, new_coords, v2p = _scene(1)
    d = lambda a: torch.from_numpy(a).cuda()
    t_xyz, t_new_xyz, t_coords, t_v2p = d(xyz), d(new_xyz), d(new_coords), d(v2p)
    idx, empty = roi_pool.voxel_query(max_range, radius, nsample, t_xyz, t_new_xyz, t_coords, t_v2p)
    # the reference kernel + the two lines of VoxelQuery.forward around it (voxel_query_utils.py:32-39)
    M = new_coords.shape[0]
    ridx = torch.zeros((M, nsample), dtype=torch.int32, device='cuda')
    torch.cuda.synchronize()
    lib.ref_voxel_query(M, v2p.shape[1], v2p.shape[2], v2p.shape[3], nsample, ctypes.c_float(radius), *max_range,
                        _p(t_new_xyz), _p(t_xyz), _p(t_coords), _p(t_v2p), _p(ridx))
    torch.cuda.synchronize()
    rempty = ridx[:, 0] == -1
    ridx[rempty] = 0
    assert torch.equal(empty, rempty)
    assert torch.equal(idx, ridx)
    assert 0 < int(empty.sum()) < M or max_range == (4, 4, 4)
    # grouping forward / backward through the module path of the reference (VoxelQueryAndGrouping.forward :80-99)
    feats = torch.randn(xyz.shape[0], 32, device='cuda', requires_grad=True)
    mod = roi_pool.VoxelQueryAndGrouping(max_range, radius, nsample)
    gf, gx, em = mod(t_coords, t_xyz, d(xyz_cnt), t_new_xyz, d(new_cnt), feats, t_v2p)
    starts = np.concatenate([[0], np.cumsum(xyz_cnt)[:-1]]).astype(np.int32)
    lidx = (ridx.view(len(xyz_cnt), -1, nsample) - d(starts).view(-1, 1, 1)).view(-1, nsample)
    lidx[rempty] = 0
    lidx = lidx.contiguous()
    rgf = torch.empty((M, 32, nsample), device='cuda')
    rgx = torch.empty((M, 3, nsample), device='cuda')
    fc = feats.detach().contiguous()
    cnt_f, cnt_q = d(xyz_cnt), d(new_cnt)          # kept alive: the raw pointers below must not dangle
    torch.cuda.synchronize()
    lib.ref_group_points(len(xyz_cnt), M, 32, nsample, _p(fc), _p(cnt_f), _p(lidx), _p(cnt_q), _p(rgf))
    lib.ref_group_points(len(xyz_cnt), M, 3, nsample, _p(t_xyz), _p(cnt_f), _p(lidx), _p(cnt_q), _p(rgx))
    torch.cuda.synchronize()
    assert torch.equal(em, rempty) and torch.equal(gf, rgf) and torch.equal(gx, rgx)
    go = torch.randn_like(gf)
    gf.backward(go)
    rg = torch.zeros_like(fc)
    go_c = go.contiguous()
    torch.cuda.synchronize()
    lib.ref_group_points_grad(len(xyz_cnt), M, 32, xyz.shape[0], nsample, _p(go_c), _p(lidx), _p(cnt_q), _p(cnt_f), _p(rg))
    torch.cuda.synchronize()
    # float atomics in a different order; row 0 of every sample collects the gradient of all empty balls (thousands of terms)
    assert float((feats.grad - rg).abs().max()) <= 1e-5 * float(rg.abs().max())


def test_voxel_query_and_grouping_vs_numpy_restatement(lib_built):
    """Small ragged case incl. empty balls, out-of-grid neighbourhoods and an empty sample, against oracle/pointnet2.py."""
    from virconv_b200 import roi_pool
    xyz, xyz_cnt, new_xyz, new_cnt, new_coords, v2p = _scene(2, n_per=(300, 1, 260), shape=(5, 40, 36), n_query=(70, 70, 70),
                                                              jitter=0.9)
    new_coords[::9, 1:] += 30                                           # some query cells far outside the grid
    d = lambda a: torch.from_numpy(a).cuda()
    for max_range, radius, nsample in (((2, 2, 2), 0.5, 8), ((4, 4, 4), 2.0, 16), ((0, 0, 0), 0.3, 4)):
        idx, empty = roi_pool.voxel_query(max_range, radius, nsample, d(xyz), d(new_xyz), d(new_coords), d(v2p))
        widx, wempty = o_pn.voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, v2p)
        assert np.array_equal(empty.cpu().numpy(), wempty) and np.array_equal(idx.cpu().numpy(), widx), max_range
        assert wempty.any() and not wempty.all()
    starts = np.concatenate([[0], np.cumsum(xyz_cnt)[:-1]])
    lidx = (widx.reshape(3, -1, nsample) - starts.reshape(-1, 1, 1)).reshape(-1, nsample).astype(np.int32)
    lidx[wempty] = 0
    feats = np.random.default_rng(0).normal(size=(xyz.shape[0], 24)).astype(np.float32)
    t_f = d(feats).requires_grad_(True)
    out = roi_pool.grouping_operation(t_f, d(xyz_cnt), d(lidx), d(new_cnt))
    assert np.array_equal(out.detach().cpu().numpy(), o_pn.group_points(feats, xyz_cnt, lidx, new_cnt))
    go = np.random.default_rng(1).normal(size=out.shape).astype(np.float32)
    out.backward(d(go))
    assert np.allclose(t_f.grad.cpu().numpy(), o_pn.group_points_grad(go, lidx, new_cnt, xyz_cnt, xyz.shape[0]), rtol=1e-5,
                       atol=1e-5)


def test_roi_pool_rejects_cpu_tensors(lib_built):
    from virconv_b200 import roi_pool
    with pytest.raises(Exception):
        roi_pool.voxel_query((1, 1, 1), 1.0, 4, torch.zeros(4, 3), torch.zeros(2, 3), torch.zeros(2, 4, dtype=torch.int32),
                             torch.zeros(1, 2, 2, 2, dtype=torch.int32))
