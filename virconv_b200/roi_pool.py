"""Voxel-RoI pooling front end: the consumers of the backbone's `x_conv3` / `x_conv4`, with the
reference's names, arguments and results, on the sm_100a kernels of csrc/roi_pool.cu.

Mirrors `pcdet/ops/pointnet2/pointnet2_stack/voxel_query_utils.py` (`voxel_query` :10-46, `VoxelQueryAndGrouping` :49-100)
and `pointnet2_utils.py` (`grouping_operation` :48-105); `generate_voxel2pinds` (pcdet/utils/spconv_utils.py:13-21) is
`ops.voxel2pinds`.  CUDA tensors only."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check


def _i32(t):
    return t if (t.dtype == torch.int32 and t.is_contiguous()) else t.to(torch.int32).contiguous()


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """-> (idx [M, nsample] int32 — GLOBAL row ids as stored in `point_indices`, rows of empty balls zeroed —,
    empty_ball_mask [M] bool).  Arguments as `VoxelQuery.forward` (voxel_query_utils.py:13-25)."""
    ops._require_cuda(xyz, new_xyz, new_coords, point_indices)
    lib = _lib.load()
    assert new_xyz.is_contiguous() and xyz.is_contiguous() and new_coords.is_contiguous() and point_indices.is_contiguous()
    assert xyz.dtype == torch.float32 and new_xyz.dtype == torch.float32
    new_coords, point_indices = _i32(new_coords), _i32(point_indices)
    M = new_coords.shape[0]
    B, Z, Y, X = point_indices.shape
    idx = torch.empty((M, nsample), dtype=torch.int32, device=xyz.device)
    empty = torch.empty((M,), dtype=torch.uint8, device=xyz.device)
    z_range, y_range, x_range = max_range
    check(lib.vc_voxel_query(M, Z, Y, X, int(nsample), float(radius), int(z_range), int(y_range), int(x_range),
                             new_xyz.data_ptr(), xyz.data_ptr(), new_coords.data_ptr(), point_indices.data_ptr(),
                             idx.data_ptr(), empty.data_ptr(), ops._stream()), 'vc_voxel_query')
    return idx, empty.bool()


class GroupingOperation(torch.autograd.Function):
    """`GroupingOperation` (pointnet2_utils.py:48-102): features [N, C], features_batch_cnt [B], idx [M, nsample]
    (batch-local rows), idx_batch_cnt [B] -> [M, C, nsample]."""

    @staticmethod
    def forward(ctx, features, features_batch_cnt, idx, idx_batch_cnt):
        ops._require_cuda(features, features_batch_cnt, idx, idx_batch_cnt)
        lib = _lib.load()
        features = features.contiguous()
        features_batch_cnt, idx, idx_batch_cnt = _i32(features_batch_cnt), _i32(idx), _i32(idx_batch_cnt)
        M, nsample = idx.shape
        N, C = features.shape
        B = idx_batch_cnt.shape[0]
        out = torch.empty((M, C, nsample), dtype=torch.float32, device=features.device)
        check(lib.vc_group_points(B, M, C, nsample, features.data_ptr(), features_batch_cnt.data_ptr(), idx.data_ptr(),
                                  idx_batch_cnt.data_ptr(), out.data_ptr(), ops._stream()), 'vc_group_points')
        ctx.for_backwards = (B, N, idx, features_batch_cnt, idx_batch_cnt)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        B, N, idx, features_batch_cnt, idx_batch_cnt = ctx.for_backwards
        M, C, nsample = grad_out.shape
        grad_out = grad_out.contiguous()
        grad_features = torch.zeros((N, C), dtype=torch.float32, device=grad_out.device)
        check(lib.vc_group_points_grad(B, M, C, N, nsample, grad_out.data_ptr(), idx.data_ptr(), idx_batch_cnt.data_ptr(),
                                       features_batch_cnt.data_ptr(), grad_features.data_ptr(), ops._stream()),
              'vc_group_points_grad')
        return grad_features, None, None, None


grouping_operation = GroupingOperation.apply


class VoxelQueryAndGrouping(nn.Module):
    """`VoxelQueryAndGrouping` (voxel_query_utils.py:49-100), same forward signature and returns."""

    def __init__(self, max_range, radius, nsample):
        super().__init__()
        self.max_range, self.radius, self.nsample = max_range, radius, nsample

    def forward(self, new_coords, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features, voxel2point_indices):
        assert xyz.shape[0] == int(xyz_batch_cnt.sum()), 'xyz: %s, xyz_batch_cnt: %s' % (str(xyz.shape), str(xyz_batch_cnt))
        assert new_coords.shape[0] == int(new_xyz_batch_cnt.sum()), \
            'new_coords: %s, new_xyz_batch_cnt: %s' % (str(new_coords.shape), str(new_xyz_batch_cnt))
        batch_size = xyz_batch_cnt.shape[0]
        idx, empty_ball_mask = voxel_query(self.max_range, self.radius, self.nsample, xyz, new_xyz, new_coords,
                                           voxel2point_indices)
        # global row ids -> batch-local ones (:85-91; the reference loops over the batch, one vectorised op here)
        starts = (torch.cumsum(xyz_batch_cnt, 0) - xyz_batch_cnt).to(torch.int32)
        idx = (idx.view(batch_size, -1, self.nsample) - starts.view(-1, 1, 1)).view(-1, self.nsample)
        idx[empty_ball_mask] = 0
        idx = idx.contiguous()
        grouped_xyz = grouping_operation(xyz, xyz_batch_cnt, idx, new_xyz_batch_cnt)
        grouped_features = grouping_operation(features, xyz_batch_cnt, idx, new_xyz_batch_cnt)
        return grouped_features, grouped_xyz, empty_ball_mask
