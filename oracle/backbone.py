"""VirConv backbones restated on the CPU oracle operators.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows `pcdet/models/backbones_3d/spconv_backbone.py`:
  NRConvBlock :150-229, post_act_block :86-107, post_act_block2d :110-131, layer_voxel_discard :134-147,
  VirConvL8x :538-699.  Module/attribute names are the reference's, so a state_dict moves between the
  reference, this oracle and `virconv_b200.backbone` unchanged (SURVEY §5 checkpoint row).

`discard_mode`: 'spconv2_compat' = what the reference does under spconv 2.x (layer discard is a silent
no-op, SURVEY §5 note); 'paper' = real order-preserving discard with the kept-row list supplied by
the caller (host RNG in the reference, `spconv_backbone.py:143`).
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import spconv_cpu as spconv
from .index2uv import index2uv


def conv_bn_relu(cin, cout, ksize, ndim, conv_type, norm_fn, stride=1, padding=0, indice_key=None):
    if conv_type == 'subm':
        cls = spconv.SubMConv3d if ndim == 3 else spconv.SubMConv2d
        conv = cls(cin, cout, ksize, bias=False, indice_key=indice_key)
    else:
        cls = spconv.SparseConv3d if ndim == 3 else spconv.SparseConv2d
        conv = cls(cin, cout, ksize, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    return spconv.SparseSequential(conv, norm_fn(cout), nn.ReLU())


class NRConvBlock(nn.Module):
    def __init__(self, input_c, output_c, stride=1, padding=1, indice_key='vir1'):
        super().__init__()
        self.stride = stride
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        c1 = input_c
        if stride > 1:
            self.down_layer = conv_bn_relu(input_c, output_c, 3, 3, 'spconv', norm_fn, stride, padding,
                                           'sp' + indice_key)
            c1 = output_c
        h = output_c // 2
        self.d3_conv1 = conv_bn_relu(c1, h, 3, 3, 'subm', norm_fn, indice_key='subm1' + indice_key)
        self.d2_conv1 = conv_bn_relu(h, h, 3, 2, 'subm', norm_fn, indice_key='subm3' + indice_key)
        self.d3_conv2 = conv_bn_relu(h, h, 3, 3, 'subm', norm_fn, indice_key='subm2' + indice_key)
        self.d2_conv2 = conv_bn_relu(h, h, 3, 2, 'subm', norm_fn, indice_key='subm4' + indice_key)

    def forward(self, x, batch_size, calib, stride, trans_param, trace=None):
        if self.stride > 1:
            x = self.down_layer(x)
        a = self.d3_conv1(x)
        b = self.d3_conv2(a)
        tp = None if trans_param is None else np.asarray(
            trans_param.detach().cpu().numpy() if torch.is_tensor(trans_param) else trans_param)
        uv = index2uv(b.indices.numpy(), batch_size, calib, stride, tp)
        img = spconv.SparseConvTensor(b.features, torch.from_numpy(uv).int(), [1600, 600], batch_size)
        c = self.d2_conv1(img)
        d = self.d2_conv2(c)
        if trace is not None:
            trace.append(dict(uv=uv, d3=b, d2=d, d3_first=a, d2_first=c, img=img))
        return b.replace_feature(torch.cat([b.features, d.features], -1))


def discard_rows(t, keep_rows):
    """Order-preserving subsample (paper-mode StVD; SURVEY §5 note)."""
    keep = torch.as_tensor(np.sort(np.asarray(keep_rows)), dtype=torch.long)
    out = spconv.SparseConvTensor(t.features[keep], t.indices[keep], t.spatial_shape, t.batch_size)
    return out


class VirConvL8x(nn.Module):
    def __init__(self, num_filters=(16, 32, 64, 64), out_features=64, input_channels=8,
                 grid_size=(1408, 1600, 80), layer_discard_rate=0.1, last_pad=0, discard_mode='spconv2_compat'):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        gs = list(grid_size)
        self.sparse_shape = [gs[2] + 1, gs[1], gs[0]]
        f = num_filters
        self.vir_conv1 = NRConvBlock(input_channels, f[0], stride=1, indice_key='vir1')
        self.vir_conv2 = NRConvBlock(f[0], f[1], stride=2, indice_key='vir2')
        self.vir_conv3 = NRConvBlock(f[1], f[2], stride=2, indice_key='vir3')
        self.vir_conv4 = NRConvBlock(f[2], f[3], stride=2, padding=(0, 1, 1), indice_key='vir4')
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(f[3], out_features, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key='spconv_down2'),
            norm_fn(out_features), nn.ReLU())
        self.layer_discard_rate = layer_discard_rate
        self.discard_mode = discard_mode

    def forward(self, voxel_features, voxel_coords, batch_size, calib, trans_param=None, keep_rows=None,
                trace=None):
        """keep_rows: optional list of 3 row-index arrays (paper-mode discard after vir_conv1..3)."""
        feats = voxel_features.clone()
        feats[:, 4:7] = 0                                   # :636
        x = spconv.SparseConvTensor(feats, voxel_coords.int(), self.sparse_shape, batch_size)
        outs = {}
        blocks = [(self.vir_conv1, 1), (self.vir_conv2, 2), (self.vir_conv3, 4), (self.vir_conv4, 8)]
        for li, (blk, stride) in enumerate(blocks):
            x = blk(x, batch_size, calib, stride, trans_param, trace)
            if self.training and self.discard_mode == 'paper' and li < 3 and keep_rows is not None:
                x = discard_rows(x, keep_rows[li])
            outs['x_conv%d' % (li + 1)] = x
        outs['out'] = self.conv_out(x)
        return outs


def decompose_tensor(t, i, batch_size):
    """spconv_backbone.py:314-337: strict `begin < x < end` (drops the x == begin column), q = shape[2] // 4."""
    q = t.spatial_shape[2] // 4
    x = t.indices[:, 3]
    mask = (i * q < x) & (x < (i + 1) * q)
    idx = t.indices[mask].clone()
    idx[:, 3] -= i * q
    return spconv.SparseConvTensor(t.features[mask], idx.int(), [t.spatial_shape[0], t.spatial_shape[1], q], batch_size)


class VirConv8x(nn.Module):
    """VirConv-T / -S backbone: LiDAR stream (shared rulebooks per stage) + virtual-point stream of NRConv blocks
    (spconv_backbone.py:232-535)."""

    def __init__(self, num_filters=(16, 32, 64, 64), out_features=64, input_channels=8, grid_size=(1408, 1600, 80),
                 layer_discard_rate=0.15, last_pad=0, mm=True, discard_mode='spconv2_compat'):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        gs = list(grid_size)
        self.sparse_shape = [gs[2] + 1, gs[1], gs[0]]
        f = num_filters
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, f[0], 3, padding=1, bias=False, indice_key='subm1'), norm_fn(f[0]), nn.ReLU())
        blk = lambda ci, co, key, **kw: conv_bn_relu(ci, co, 3, 3, kw.pop('conv_type', 'subm'), norm_fn, indice_key=key, **kw)
        self.conv1 = spconv.SparseSequential(blk(f[0], f[0], 'subm1'))
        self.conv2 = spconv.SparseSequential(blk(f[0], f[1], 'spconv2', conv_type='spconv', stride=2, padding=1),
                                             blk(f[1], f[1], 'subm2'), blk(f[1], f[1], 'subm2'))
        self.conv3 = spconv.SparseSequential(blk(f[1], f[2], 'spconv3', conv_type='spconv', stride=2, padding=1),
                                             blk(f[2], f[2], 'subm3'), blk(f[2], f[2], 'subm3'))
        self.conv4 = spconv.SparseSequential(blk(f[2], f[3], 'spconv4', conv_type='spconv', stride=2, padding=(0, 1, 1)),
                                             blk(f[3], f[3], 'subm4'), blk(f[3], f[3], 'subm4'))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(f[3], out_features, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key='spconv_down2'), norm_fn(out_features), nn.ReLU())
        self.mm = mm
        if mm:
            self.vir_conv1 = NRConvBlock(input_channels, f[0], stride=1, indice_key='vir1')
            self.vir_conv2 = NRConvBlock(f[0], f[1], stride=2, indice_key='vir2')
            self.vir_conv3 = NRConvBlock(f[1], f[2], stride=2, indice_key='vir3')
            self.vir_conv4 = NRConvBlock(f[2], f[3], stride=2, padding=(0, 1, 1), indice_key='vir4')
        self.layer_discard_rate = layer_discard_rate
        self.discard_mode = discard_mode

    def _lidar(self, feats, coords, shape, batch_size):
        x = spconv.SparseConvTensor(feats, coords.int(), shape, batch_size)
        x1 = self.conv1(self.conv_input(x))
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        return x1, x2, x3, x4, self.conv_out(x4)

    def forward(self, arrays, batch_size, calib, aug_param=None, transform_param=None, keep_rows=None):
        """arrays: dict of torch tensors keyed like batch_dict (voxel_features{,i}, voxel_coords{,i}, *_mm{,i}).
        Returns a dict with the reference's output keys."""
        rot_num = 1 if transform_param is None else transform_param.shape[1]
        sfx = [''] + [str(i) for i in range(1, rot_num)]
        out = {}
        if self.training:
            for s in sfx:
                x1, x2, x3, x4, o = self._lidar(arrays['voxel_features' + s], arrays['voxel_coords' + s],
                                                self.sparse_shape, batch_size)
                out['encoded_spconv_tensor' + s] = o
                out['multi_scale_3d_features' + s] = dict(x_conv1=x1, x_conv2=x2, x_conv3=x3, x_conv4=x4)
        else:
            feats, coords = [], []
            for i, s in enumerate(sfx):
                feats.append(arrays['voxel_features' + s])
                c = arrays['voxel_coords' + s].clone()
                c[:, 3] += i * self.sparse_shape[2]
                coords.append(c)
            big = [self.sparse_shape[0], self.sparse_shape[1], self.sparse_shape[2] * 4]
            x1, x2, x3, x4, o = self._lidar(torch.cat(feats, 0), torch.cat(coords), big, batch_size)
            for i, s in enumerate(sfx):
                out['encoded_spconv_tensor' + s] = decompose_tensor(o, i, batch_size)
                out['multi_scale_3d_features' + s] = dict(x_conv1=None, x_conv2=None,
                                                          x_conv3=decompose_tensor(x3, i, batch_size),
                                                          x_conv4=decompose_tensor(x4, i, batch_size))
        if self.mm:
            for i, s in enumerate(sfx):
                tp = aug_param
                if transform_param is not None:
                    tp = transform_param[:, i, :]
                x = spconv.SparseConvTensor(arrays['voxel_features_mm' + s], arrays['voxel_coords_mm' + s].int(),
                                            self.sparse_shape, batch_size)
                paper = self.training and self.discard_mode == 'paper' and keep_rows is not None
                if paper:
                    x = discard_rows(x, keep_rows[0])
                feats = {}
                blocks = [(self.vir_conv1, 1), (self.vir_conv2, 2), (self.vir_conv3, 4), (self.vir_conv4, 8)]
                for li, (blk, stride) in enumerate(blocks):
                    x = blk(x, batch_size, calib, stride, tp)
                    if paper and li < 3:
                        x = discard_rows(x, keep_rows[li + 1])
                    feats['x_conv%d' % (li + 1)] = x
                out['multi_scale_3d_features_mm' + s] = feats
        return out
