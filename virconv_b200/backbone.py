"""VirConv 3-D backbones on the virconv_b200 operators — drop-in for the classes the detector registry
instantiates (`pcdet/models/backbones_3d/__init__.py:3-8`, `detector3d_template.py:65-78`):
same constructor `(model_cfg, input_channels, grid_size, **kwargs)`, same `.forward(batch_dict)` keys,
same `.num_point_features`, same sub-module names (=> same state_dict keys).

Restated from `pcdet/models/backbones_3d/spconv_backbone.py`: NRConvBlock :150-229, VirConvL8x :538-699,
layer_voxel_discard :134-147.  Differences that do not change results:
  * conv + BatchNorm1d + ReLU run as one fused op; the d3/d2 conv pairs of a block share one rulebook
    (identical indices) although the reference builds it twice under distinct indice_keys;
  * `index2uv` is one kernel over the whole batch (the reference loops over samples with boolean-mask
    indexing, :61-78); its per-sample matrices are built once per forward;
  * StVD layer discard: `discard_mode='spconv2_compat'` reproduces the reference under spconv 2.x, where
    `layer_voxel_discard` returns None after rebinding a local and the caller's tensor is unchanged
    (SURVEY §5 note) — i.e. no discard; `discard_mode='paper'` really drops rows, order-preserving.
"""
from __future__ import annotations

from functools import partial

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import spconv_compat as spconv


def _cfg_get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default) if not hasattr(cfg, 'get') else cfg.get(key, default)


def sparse_block(cin, cout, kernel_size, norm_fn, ndim=3, conv_type='subm', stride=1, padding=0, indice_key=None):
    """conv(bias=False) -> norm -> ReLU, as post_act_block / post_act_block2d (:86-131)."""
    if conv_type == 'subm':
        cls = spconv.SubMConv3d if ndim == 3 else spconv.SubMConv2d
        conv = cls(cin, cout, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == 'spconv':
        cls = spconv.SparseConv3d if ndim == 3 else spconv.SparseConv2d
        conv = cls(cin, cout, kernel_size, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    else:
        raise NotImplementedError(conv_type)
    return spconv.SparseSequential(conv, norm_fn(cout), nn.ReLU())


class NRConvBlock(nn.Module):
    """Noise-resistant conv block: [strided down conv] -> 2x SubMConv3d -> project voxels to pixels ->
    2x SubMConv2d in image space -> concat (spconv_backbone.py:150-229)."""

    def __init__(self, input_c=16, output_c=16, stride=1, padding=1, indice_key='vir1', conv_depth=False):
        super().__init__()
        self.stride = stride
        self.conv_depth = conv_depth
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        if stride > 1:
            self.down_layer = sparse_block(input_c, output_c, 3, norm_fn, 3, 'spconv', stride, padding,
                                           'sp' + indice_key)
        c1 = output_c if stride > 1 else input_c
        if conv_depth:
            c1 += 4
        half = output_c // 2
        self.d3_conv1 = sparse_block(c1, half, 3, norm_fn, 3, 'subm', padding=1, indice_key='subm1' + indice_key)
        self.d2_conv1 = sparse_block(half, half, 3, norm_fn, 2, 'subm', padding=1, indice_key='subm3' + indice_key)
        self.d3_conv2 = sparse_block(half, half, 3, norm_fn, 3, 'subm', padding=1, indice_key='subm2' + indice_key)
        self.d2_conv2 = sparse_block(half, half, 3, norm_fn, 2, 'subm', padding=1, indice_key='subm4' + indice_key)

    def forward(self, sp_tensor, batch_size, proj_params, stride):
        if self.stride > 1:
            sp_tensor = self.down_layer(sp_tensor)
        d3 = self.d3_conv2(self.d3_conv1(sp_tensor))
        uv = ops.index2uv(spconv._as_i32(d3.indices), batch_size, proj_params, stride)
        img = spconv.SparseConvTensor(d3.features, uv, [1600, 600], batch_size)
        img.features_bf16 = d3.features_bf16
        d2 = self.d2_conv2(self.d2_conv1(img))
        cat = ops.Cat2Fn.apply(d3.features, d2.features, d3.features_bf16 is not None)
        if isinstance(cat, tuple):
            out = d3.replace_feature(cat[0])
            out.features_bf16 = cat[1]
            return out
        return d3.replace_feature(cat)


def stvd_keep_rows(n, rate, rng=np.random):
    """Rows kept by StVD layer discard: the reference draws `np.random.permutation(n)[:int(n*(1-rate))]`
    (:143-144); kept order-preserving (sorted) so published tensors stay batch-contiguous."""
    keep = rng.permutation(n)[:int(n * (1 - rate))]
    return np.sort(keep)


def discard_rows(t, keep_rows):
    rows = torch.as_tensor(np.ascontiguousarray(keep_rows), dtype=torch.int32).to(t.features.device)
    feats = ops.GatherRowsFn.apply(t.features, rows)
    idx = ops.gather_rows(spconv._as_i32(t.indices), rows)
    return spconv.SparseConvTensor(feats, idx, t.spatial_shape, t.batch_size)


class VirConvL8x(nn.Module):
    """VirConv-L backbone (fused LiDAR + virtual stream), spconv_backbone.py:538-699."""

    def __init__(self, model_cfg, input_channels, grid_size, discard_mode='spconv2_compat', precision='fp32',
                 **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.return_num_features_as_dict = _cfg_get(model_cfg, 'RETURN_NUM_FEATURES_AS_DICT', False)
        self.out_features = _cfg_get(model_cfg, 'OUT_FEATURES', 64)
        self.layer_discard_rate = _cfg_get(model_cfg, 'LAYER_DISCARD_RATE', 0.0)
        num_filters = list(_cfg_get(model_cfg, 'NUM_FILTERS', [16, 32, 64, 64]))
        assert discard_mode in ('spconv2_compat', 'paper')
        self.discard_mode = discard_mode
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        gs = [int(g) for g in grid_size]
        self.sparse_shape = [gs[2] + 1, gs[1], gs[0]]            # grid_size[::-1] + [1, 0, 0]   (:552)

        self.vir_conv1 = NRConvBlock(input_channels, num_filters[0], stride=1, indice_key='vir1')
        self.vir_conv2 = NRConvBlock(num_filters[0], num_filters[1], stride=2, indice_key='vir2')
        self.vir_conv3 = NRConvBlock(num_filters[1], num_filters[2], stride=2, indice_key='vir3')
        self.vir_conv4 = NRConvBlock(num_filters[2], num_filters[3], stride=2, padding=(0, 1, 1), indice_key='vir4')
        last_pad = _cfg_get(model_cfg, 'last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(num_filters[3], self.out_features, (3, 1, 1), stride=(2, 1, 1), padding=last_pad,
                                bias=False, indice_key='spconv_down2'),
            norm_fn(self.out_features), nn.ReLU())
        spconv.set_precision(self, precision)
        self.num_point_features = self.out_features
        if self.return_num_features_as_dict:
            self.num_point_features = {'x_conv%d' % (i + 1): num_filters[i] for i in range(4)}

    def _maybe_discard(self, t, batch_dict, layer):
        if not self.training or self.discard_mode != 'paper' or self.layer_discard_rate == 0:
            return t
        given = batch_dict.get('stvd_keep_rows')
        keep = given[layer] if given is not None else stvd_keep_rows(t.features.shape[0], self.layer_discard_rate)
        return discard_rows(t, keep)

    def forward(self, batch_dict):
        rot_num = batch_dict['transform_param'].shape[1] if 'transform_param' in batch_dict else 1
        batch_size = batch_dict['batch_size']
        calib = batch_dict['calib']
        ops.stats_arena(batch_dict['voxel_features'].device).reset()    # one memset for all BN accumulators of this pass
        for i in range(rot_num):
            rid = '' if i == 0 else str(i)
            feats, coords = batch_dict['voxel_features' + rid], batch_dict['voxel_coords' + rid]
            feats[:, 4:7] = 0                                    # RGB channels unused (:636), in place like the reference
            x = spconv.SparseConvTensor(feats, coords.int(), self.sparse_shape, batch_size)
            trans = batch_dict['aug_param'] if 'aug_param' in batch_dict else None
            if 'transform_param' in batch_dict:
                trans = batch_dict['transform_param'][:, i, :]
            proj = ops.projection_params(calib, trans, batch_size, feats.device)

            x1 = self.vir_conv1(x, batch_size, proj, 1)
            x1 = self._maybe_discard(x1, batch_dict, 0)
            x2 = self.vir_conv2(x1, batch_size, proj, 2)
            x2 = self._maybe_discard(x2, batch_dict, 1)
            x3 = self.vir_conv3(x2, batch_size, proj, 4)
            x3 = self._maybe_discard(x3, batch_dict, 2)
            x4 = self.vir_conv4(x3, batch_size, proj, 8)
            out = self.conv_out(x4)

            batch_dict.update({
                'encoded_spconv_tensor' + rid: out,
                'encoded_spconv_tensor_stride' + rid: 8,
                'multi_scale_3d_features' + rid: {'x_conv1': x1, 'x_conv2': x2, 'x_conv3': x3, 'x_conv4': x4},
                'multi_scale_3d_strides' + rid: {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8},
            })
        return batch_dict
