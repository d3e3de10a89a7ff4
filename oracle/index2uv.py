"""Voxel index -> image pixel index (`index2uv`) — CPU oracle in numpy float32.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, line by line:
  * `index2points`           pcdet/models/backbones_3d/spconv_backbone.py:8-24
  * `index2uv`               pcdet/models/backbones_3d/spconv_backbone.py:54-83
  * inverse augmentation     pcdet/datasets/augmentor/X_transform.py:139-154 (queue reversed: scaling, flip,
                             rotation, :28-35) with `world_scaling` :98-120, `random_flip_with_param`
                             augmentor_utils.py:26-32, `rotate_points_along_z` common_utils.py:34-56
  * `lidar_to_rect_cuda`     pcdet/utils/calibration_kitti.py:120-130
  * `rect_to_img_cuda`       pcdet/utils/calibration_kitti.py:143-153 (divides by rect z)

The reference evaluates the two small matmuls with torch (BLAS / cuBLAS: accumulation order and FMA
use unspecified).  The oracle fixes ONE order — left-to-right multiply then add, every operation
rounded to float32, no FMA — and the CUDA kernel implements exactly that with `__fmul_rn/__fadd_rn`,
so oracle and kernel agree bit for bit; `oracle/make_golden.py` measures how often this order
differs from the reference's torch evaluation after truncation (see tests/golden/README.md).
float -> int32 conversion follows the CUDA rule the reference's GPU path gets from `.int()`:
round toward zero, saturate, NaN -> 0.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def compose_lidar_to_rect(V2C, R0):
    """M[4,3] = V2C.T @ R0.T in float32, fixed order."""
    a = np.asarray(V2C, dtype=F).T  # [4,3]
    b = np.asarray(R0, dtype=F).T   # [3,3]
    m = np.zeros((4, 3), dtype=F)
    for i in range(4):
        for j in range(3):
            m[i, j] = F(F(F(a[i, 0] * b[0, j]) + F(a[i, 1] * b[1, j])) + F(a[i, 2] * b[2, j]))
    return m


def trunc_sat_i32(x):
    x = np.asarray(x, dtype=F)
    t = np.trunc(np.nan_to_num(x.astype(np.float64), nan=0.0, posinf=2.0 ** 40, neginf=-2.0 ** 40))
    return np.clip(t, -2147483648.0, 2147483647.0).astype(np.int64).astype(np.int32)


def index2uv(indices, batch_size, calibs, stride, trans_param=None, pts_range=(0, -40, -3, 70.4, 40, 1),
             voxel_size=(0.05, 0.05, 0.05)):
    idx = np.asarray(indices)
    vs = np.array(voxel_size, dtype=np.float64) * stride
    mn = [pts_range[0] + vs[0] / 2, pts_range[1] + vs[1] / 2, pts_range[2] + vs[2] / 2]
    x = idx[:, 3].astype(F) * F(vs[0]) + F(mn[0])
    y = idx[:, 2].astype(F) * F(vs[1]) + F(mn[1])
    z = idx[:, 1].astype(F) * F(vs[2]) + F(mn[2])
    uv = np.zeros((idx.shape[0], 3), dtype=np.int32)
    uv[:, 0] = idx[:, 0]
    for b in range(batch_size):
        m = idx[:, 0] == b
        xb, yb, zb = x[m], y[m], z[m]
        if trans_param is not None:
            rot, flip, scale = (F(v) for v in np.asarray(trans_param, dtype=F)[b])
            xb, yb, zb = xb / scale, yb / scale, zb / scale
            if flip != 0:
                yb = -yb
            c, s = np.cos(F(-rot)).astype(F), np.sin(F(-rot)).astype(F)
            xr = xb * c + yb * F(-s)
            yr = xb * s + yb * c
            xb, yb = xr, yr
        M = compose_lidar_to_rect(calibs[b].V2C, calibs[b].R0)
        P = np.asarray(calibs[b].P2, dtype=F).T  # [4,3]
        r = [((xb * M[0, j] + yb * M[1, j]) + zb * M[2, j]) + M[3, j] for j in range(3)]
        h = [((r[0] * P[0, j] + r[1] * P[1, j]) + r[2] * P[2, j]) + P[3, j] for j in range(2)]
        with np.errstate(divide='ignore', invalid='ignore'):
            u = h[0] / r[2]
            v = h[1] / r[2]
        uv[m, 1] = np.clip(trunc_sat_i32(u), 0, 1400 - 1) // stride
        uv[m, 2] = np.clip(trunc_sat_i32(v), 0, 600 - 1) // stride
    return uv
