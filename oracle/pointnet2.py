"""CPU restatement of the stacked pointnet2 ops the RoI head runs on the backbone's outputs — TEST INFRASTRUCTURE
(see oracle/__init__.py).  Plain loops, small cases only; at full size the checker is the reference's own kernels
compiled into oracle/_ref (oracle/ref_build.py).

voxel_query  : pcdet/ops/pointnet2/pointnet2_stack/src/voxel_query_gpu.cu:10-89 + voxel_query_utils.py:36-41
group_points : .../src/group_points_gpu.cu:71-103, grad :15-45"""
from __future__ import annotations

import numpy as np


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """-> (idx [M, nsample] int32 with empty rows zeroed, empty_ball_mask [M] bool)"""
    M = new_coords.shape[0]
    B, R1, R2, R3 = point_indices.shape
    zr, yr, xr = max_range
    idx = np.zeros((M, nsample), dtype=np.int32)
    f32 = np.float32
    radius2 = f32(radius) * f32(radius)
    flat = point_indices.reshape(-1)
    for pt in range(M):
        b, cz, cy, cx = (int(v) for v in new_coords[pt])
        nx, ny, nz = (f32(v) for v in new_xyz[pt])
        cnt = 0
        for dz in range(-zr, zr + 1):
            z = cz + dz
            if z < 0 or z >= R1:
                continue
            for dy in range(-yr, yr + 1):
                y = cy + dy
                if y < 0 or y >= R2:
                    continue
                for dx in range(-xr, xr + 1):
                    x = cx + dx
                    if x < 0 or x >= R3:
                        continue
                    nb = int(flat[b * R1 * R2 * R3 + z * R2 * R3 + y * R3 + x])
                    if nb < 0:
                        continue
                    d = xyz[nb].astype(np.float64) - np.array([nx, ny, nz], dtype=np.float64)
                    dist2 = f32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])      # exact products, rounded once (tests keep
                    if dist2 > radius2:                                       # their points off the radius boundary)
                        continue
                    if cnt < nsample:
                        if cnt == 0:
                            idx[pt, :] = nb
                        idx[pt, cnt] = nb
                        cnt += 1
        if cnt == 0:
            idx[pt, 0] = -1
    empty = idx[:, 0] == -1
    idx[empty] = 0
    return idx, empty


def _starts(pt, idx_batch_cnt, features_batch_cnt):
    bs, pt_cnt = 0, int(idx_batch_cnt[0])
    for k in range(1, len(idx_batch_cnt)):
        if pt < pt_cnt:
            break
        pt_cnt += int(idx_batch_cnt[k])
        bs = k
    return int(np.sum(features_batch_cnt[:bs]))


def group_points(features, features_batch_cnt, idx, idx_batch_cnt):
    M, nsample = idx.shape
    C = features.shape[1]
    out = np.zeros((M, C, nsample), dtype=np.float32)
    for pt in range(M):
        s0 = _starts(pt, idx_batch_cnt, features_batch_cnt)
        out[pt] = features[s0 + idx[pt]].T
    return out


def group_points_grad(grad_out, idx, idx_batch_cnt, features_batch_cnt, N):
    M, C, nsample = grad_out.shape
    g = np.zeros((N, C), dtype=np.float64)
    for pt in range(M):
        s0 = _starts(pt, idx_batch_cnt, features_batch_cnt)
        for s in range(nsample):
            g[s0 + idx[pt, s]] += grad_out[pt, :, s]
    return g.astype(np.float32)
