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
