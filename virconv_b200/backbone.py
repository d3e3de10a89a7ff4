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

from . import executor, ops
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
        idx = spconv._as_i32(sp_tensor.indices)
        rb2d = None
        if ops.OVERLAP and idx.shape[0] > 0:
            # projection + the image-space rulebook only need the indices: build them on the side stream while the two
            # 3-D convs run on the main one
            sd = ops.side(idx.device)
            main = torch.cuda.current_stream()
            sd.stream.wait_stream(main)
            with torch.cuda.stream(sd.stream):
                uv = ops.index2uv(idx, batch_size, proj_params, stride)
                conv2d = self.d2_conv1[0]
                rb2d = ops.build_subm_rulebook(uv, batch_size, [1600, 600], conv2d.kernel_size, conv2d.dilation)
            ops._keep_alive_on(sd.stream, idx, proj_params)
            ops._keep_alive_on(main, uv, rb2d.nbr, rb2d.pair_num)
        d3 = self.d3_conv2(self.d3_conv1(sp_tensor))
        if rb2d is None:
            uv = ops.index2uv(spconv._as_i32(d3.indices), batch_size, proj_params, stride)
        else:
            torch.cuda.current_stream().wait_stream(sd.stream)
        img = spconv.SparseConvTensor(d3.features, uv, [1600, 600], batch_size)
        if rb2d is not None:
            rb2d._keepalive = uv
            img.indice_dict[self.d2_conv1[0].indice_key] = rb2d       # both image convs see identical indices
            img.indice_dict[self.d2_conv2[0].indice_key] = rb2d
        img.features_bf16 = d3.features_bf16
        d2 = self.d2_conv2(self.d2_conv1(img))
        cat = ops.Cat2Fn.apply(d3.features, d2.features, d3.features_bf16 is not None)
        if isinstance(cat, tuple):
            out = d3.replace_feature(cat[0])
            out.features_bf16 = cat[1]
            return out
        return d3.replace_feature(cat)


def plan_nrconv(plan, block, f, iset, proj_stride):
    """Record one NRConvBlock (NRConvBlock.forward above; spconv_backbone.py:207-229) into an executor plan.  Index ops
    first (they go to the side stream), then the feature ops.  -> (feature slot of the concat, index set of the block)."""
    if block.stride > 1:
        conv = block.down_layer[0]
        iset, rb = plan.conv_rb(iset, 3, conv.kernel_size, conv.stride, conv.padding, conv.dilation, keys=[conv.indice_key])
        f = plan.cbr(f, rb, conv, block.down_layer[1], seq=block.down_layer)
    c3, c2 = block.d3_conv1[0], block.d2_conv1[0]
    assert (block.d3_conv2[0].kernel_size, block.d3_conv2[0].dilation) == (c3.kernel_size, c3.dilation)
    assert (block.d2_conv2[0].kernel_size, block.d2_conv2[0].dilation) == (c2.kernel_size, c2.dilation)
    rb3 = plan.subm_rb(iset, 3, c3.kernel_size, c3.dilation, unique=True, keys=[c3.indice_key, block.d3_conv2[0].indice_key],
                       stream=2)        # stream 1 keeps the strided convs' count -> indices -> tables chain to itself
    uv = plan.index2uv(iset, proj_stride, stream=2)
    rb2 = plan.subm_rb(uv, 2, c2.kernel_size, c2.dilation, unique=False,      # projected pixels collide
                       keys=[c2.indice_key, block.d2_conv2[0].indice_key], stream=2)
    d3 = plan.cbr(plan.cbr(f, rb3, c3, block.d3_conv1[1], seq=block.d3_conv1), rb3, block.d3_conv2[0], block.d3_conv2[1],
                  seq=block.d3_conv2)
    d2 = plan.cbr(plan.cbr(d3, rb2, c2, block.d2_conv1[1], seq=block.d2_conv1), rb2, block.d2_conv2[0], block.d2_conv2[1],
                  seq=block.d2_conv2)
    return plan.cat(d3, d2), iset


def _published_tensor(res, name, batch_size, indice_dict):
    f, idx, shape, fb, cnt = res[name]
    t = spconv.SparseConvTensor(f, idx, shape, batch_size, indice_dict=indice_dict)
    t.features_bf16 = fb
    t.num_rows = cnt          # static mode: device int32[1] row count (features / indices are capacity sized, zero / -1 tail)
    return t


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

    # -- native plan (csrc/executor.cu): the whole forward / backward in one C-ABI call each -------------------
    def _plan(self):
        if getattr(self, '_plan_cache', None) is None:
            plan = executor.Plan(self.vir_conv1.d3_conv1[0].in_channels)
            f, iset = 0, 0
            for i, (blk, s) in enumerate(zip((self.vir_conv1, self.vir_conv2, self.vir_conv3, self.vir_conv4), (1, 2, 4, 8))):
                f, iset = plan_nrconv(plan, blk, f, iset, s)
                plan.publish('x_conv%d' % (i + 1), f, iset)
            conv = self.conv_out[0]
            iset, rb = plan.conv_rb(iset, 3, conv.kernel_size, conv.stride, conv.padding, conv.dilation, keys=[conv.indice_key])
            plan.publish('out', plan.cbr(f, rb, conv, self.conv_out[1], seq=self.conv_out), iset)
            self._plan_cache = plan
        return self._plan_cache

    def _use_plan(self, feats):
        if not executor.ENABLED or not feats.is_cuda or feats.shape[0] == 0 or feats.requires_grad:
            return False             # (the plan does not propagate a gradient into the voxel features; the module path does)
        if self.training and self.discard_mode == 'paper' and self.layer_discard_rate != 0:
            return False             # row-dropping StVD between the blocks is only on the module path
        plan = self._plan()
        if not plan.is_current():          # a sub-module was replaced after the plan was built (e.g. SyncBatchNorm conversion)
            self._plan_cache = None
            plan = self._plan()
        if not plan.eligible():            # (checked every forward: dtype / layout / module type may change under .half(), .to())
            return False
        mode = plan.layers[0][1].training
        return all(bn.training == mode for _, bn in plan.layers)

    def index_phase(self, batch_dict):
        """Phased static execution (graph.PipelinedStep), phase 1: everything of the forward that depends on the voxel
        COORDINATES only — int32 coordinates, rulebooks of every stage, voxel -> pixel projections — enqueued into
        batch_dict['virconv_static'].arena.  forward() with the same batch_dict and `static.phase = 2` then runs the
        feature operators over it."""
        feats, coords = batch_dict['voxel_features'], batch_dict['voxel_coords']
        static = batch_dict['virconv_static']
        assert static.phase == 1 and self._use_plan(feats)
        ci = spconv._as_i32(coords)
        executor.run_index(self._plan(), feats, ci, self.sparse_shape, batch_dict['batch_size'], batch_dict['virconv_proj'],
                           self.conv_out[1].training, self.conv_out[0].precision, static)
        batch_dict['virconv_ci'] = ci
        return batch_dict

    def forward(self, batch_dict):
        rot_num = batch_dict['transform_param'].shape[1] if 'transform_param' in batch_dict else 1
        batch_size = batch_dict['batch_size']
        calib = batch_dict['calib']
        ops.stats_arena(batch_dict['voxel_features'].device).reset()    # one memset for all BN accumulators of this pass
        for i in range(rot_num):
            rid = '' if i == 0 else str(i)
            feats, coords = batch_dict['voxel_features' + rid], batch_dict['voxel_coords' + rid]
            feats[:, 4:7] = 0                                    # RGB channels unused (:636), in place like the reference
            trans = batch_dict['aug_param'] if 'aug_param' in batch_dict else None
            if 'transform_param' in batch_dict:
                trans = batch_dict['transform_param'][:, i, :]
            if self._use_plan(feats):
                bn_training = self.conv_out[1].training
                # 'virconv_inputs_ready': the caller's promise that voxel_coords is resident and complete (a prefetched
                # batch) -> the index pipeline of this step (starting with the .int() of the coordinates, :641) runs on
                # the side stream without waiting for main, i.e. it may overlap the previous step's backward
                side = ops.side(feats.device).stream if executor.TWO_STREAMS else None
                static = batch_dict.get('virconv_static')
                ready = side is not None and static is None and bool(batch_dict.get('virconv_inputs_ready', False))
                if static is not None and static.phase == 2:
                    ci = batch_dict['virconv_ci']           # (built by index_phase for this very batch)
                elif ready and (coords.dtype != torch.int32 or not coords.is_contiguous()):
                    executor.reserve_blocks('coords_i32', 4 * coords.numel(), feats.device, side)
                    with torch.cuda.stream(side):
                        ci = spconv._as_i32(coords)
                    ci.record_stream(torch.cuda.current_stream(feats.device))   # published as x_conv1's indices
                else:
                    ci = spconv._as_i32(coords)
                # static mode ('virconv_static': executor.StaticSpec): capacity-sized inputs + device row count, projection
                # block already on the device ('virconv_proj', [B,28], ops.projection_params) — nothing below reads a
                # device value or uploads from pageable memory, so the call can be captured into a CUDA graph
                proj = batch_dict.get('virconv_proj')
                if proj is None:
                    assert static is None, "static mode needs batch_dict['virconv_proj'] (ops.projection_params on the device)"
                    proj = ops.projection_params(calib, trans, batch_size, feats.device, side)
                run, res = executor.run_plan(self._plan(), feats, ci, self.sparse_shape, batch_size, proj, bn_training,
                                             self.conv_out[0].precision, inputs_ready=ready, static=static)
                executor.note_last_run(self, run)   # (graph.GraphedStep sizes its capacities from an exact-mode run)
                idict = executor.LazyIndiceDict(run, ci, self.sparse_shape)
                x1, x2, x3, x4, out = (_published_tensor(res, k, batch_size, idict)
                                       for k in ('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4', 'out'))
                batch_dict.update({
                    'encoded_spconv_tensor' + rid: out,
                    'encoded_spconv_tensor_stride' + rid: 8,
                    'multi_scale_3d_features' + rid: {'x_conv1': x1, 'x_conv2': x2, 'x_conv3': x3, 'x_conv4': x4},
                    'multi_scale_3d_strides' + rid: {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8},
                })
                continue
            proj = ops.projection_params(calib, trans, batch_size, feats.device)
            x = spconv.SparseConvTensor(feats, coords.int(), self.sparse_shape, batch_size)

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


def decompose_tensor(t, i, batch_size):
    """Split the i-th transformed copy back out of an x-axis-batched tensor (spconv_backbone.py:314-337):
    rows with `i*q < x < (i+1)*q` (strict on both sides — the reference drops the x == i*q column), q = W // 4."""
    q = t.spatial_shape[2] // 4
    idx = spconv._as_i32(t.indices)
    x = idx[:, 3]
    rows = torch.nonzero((x > i * q) & (x < (i + 1) * q)).squeeze(1).to(torch.int32)   # data-dependent size (eval only)
    feats = ops.gather_rows(t.features, rows)
    sub = ops.gather_rows(idx, rows)
    sub[:, 3] -= i * q
    return spconv.SparseConvTensor(feats, sub, [t.spatial_shape[0], t.spatial_shape[1], q], batch_size)


class VirConv8x(nn.Module):
    """VirConv-T / VirConv-S backbone (spconv_backbone.py:232-535): a LiDAR stream whose submanifold convs share one
    rulebook per stage (indice_keys 'subm1'..'subm4') plus, when `MM`, the virtual-point stream of NRConv blocks.
    Eval mode batches the ROT_NUM transformed copies along x into one [D, H, 4W] tensor (:360,:418-432) and splits the
    published tensors afterwards."""

    def __init__(self, model_cfg, input_channels, grid_size, discard_mode='spconv2_compat', precision='fp32', **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.return_num_features_as_dict = _cfg_get(model_cfg, 'RETURN_NUM_FEATURES_AS_DICT', False)
        self.out_features = _cfg_get(model_cfg, 'OUT_FEATURES', 64)
        self.layer_discard_rate = _cfg_get(model_cfg, 'LAYER_DISCARD_RATE', 0.0)
        f = list(_cfg_get(model_cfg, 'NUM_FILTERS', [16, 32, 64, 64]))
        self.mm = bool(_cfg_get(model_cfg, 'MM', False))
        assert discard_mode in ('spconv2_compat', 'paper')
        self.discard_mode = discard_mode
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        gs = [int(g) for g in grid_size]
        self.sparse_shape = [gs[2] + 1, gs[1], gs[0]]

        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, f[0], 3, padding=1, bias=False, indice_key='subm1'), norm_fn(f[0]), nn.ReLU())
        sub = lambda ci, co, key: sparse_block(ci, co, 3, norm_fn, 3, 'subm', padding=1, indice_key=key)
        down = lambda ci, co, key, pad: sparse_block(ci, co, 3, norm_fn, 3, 'spconv', stride=2, padding=pad, indice_key=key)
        self.conv1 = spconv.SparseSequential(sub(f[0], f[0], 'subm1'))
        self.conv2 = spconv.SparseSequential(down(f[0], f[1], 'spconv2', 1), sub(f[1], f[1], 'subm2'), sub(f[1], f[1], 'subm2'))
        self.conv3 = spconv.SparseSequential(down(f[1], f[2], 'spconv3', 1), sub(f[2], f[2], 'subm3'), sub(f[2], f[2], 'subm3'))
        self.conv4 = spconv.SparseSequential(down(f[2], f[3], 'spconv4', (0, 1, 1)), sub(f[3], f[3], 'subm4'),
                                             sub(f[3], f[3], 'subm4'))
        last_pad = _cfg_get(model_cfg, 'last_pad', 0)
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(f[3], self.out_features, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key='spconv_down2'), norm_fn(self.out_features), nn.ReLU())
        if self.mm:
            self.vir_conv1 = NRConvBlock(input_channels, f[0], stride=1, indice_key='vir1')
            self.vir_conv2 = NRConvBlock(f[0], f[1], stride=2, indice_key='vir2')
            self.vir_conv3 = NRConvBlock(f[1], f[2], stride=2, indice_key='vir3')
            self.vir_conv4 = NRConvBlock(f[2], f[3], stride=2, padding=(0, 1, 1), indice_key='vir4')
        spconv.set_precision(self, precision)
        self.num_point_features = self.out_features
        if self.return_num_features_as_dict:
            self.num_point_features = {'x_conv%d' % (i + 1): f[i] for i in range(4)}

    # -- native plans (csrc/executor.cu) -----------------------------------------------------------------------
    def _plan_lidar(self):
        """LiDAR stream (:248-291): the submanifold convs of a stage share one rulebook ('subm1'..'subm4')."""
        if getattr(self, '_plan_lidar_cache', None) is None:
            plan = executor.Plan(self.conv_input[0].in_channels)
            c0 = self.conv_input[0]
            rb = plan.subm_rb(0, 3, c0.kernel_size, c0.dilation, unique=True, keys=[c0.indice_key])
            f = plan.cbr(0, rb, c0, self.conv_input[1], seq=self.conv_input)
            f = plan.cbr(f, rb, self.conv1[0][0], self.conv1[0][1], seq=self.conv1[0])
            iset = 0
            plan.publish('x_conv1', f, iset)
            for name, stage in (('x_conv2', self.conv2), ('x_conv3', self.conv3), ('x_conv4', self.conv4)):
                down = stage[0][0]
                iset, rbd = plan.conv_rb(iset, 3, down.kernel_size, down.stride, down.padding, down.dilation, keys=[down.indice_key])
                c1 = stage[1][0]
                rb = plan.subm_rb(iset, 3, c1.kernel_size, c1.dilation, unique=True, keys=[c1.indice_key])
                f = plan.cbr(f, rbd, down, stage[0][1], seq=stage[0])
                f = plan.cbr(f, rb, c1, stage[1][1], seq=stage[1])
                f = plan.cbr(f, rb, stage[2][0], stage[2][1], seq=stage[2])
                plan.publish(name, f, iset)
            co = self.conv_out[0]
            iset, rbo = plan.conv_rb(iset, 3, co.kernel_size, co.stride, co.padding, co.dilation, keys=[co.indice_key])
            plan.publish('out', plan.cbr(f, rbo, co, self.conv_out[1], seq=self.conv_out), iset)
            self._plan_lidar_cache = plan
        return self._plan_lidar_cache

    def _plan_mm(self):
        if getattr(self, '_plan_mm_cache', None) is None:
            plan = executor.Plan(self.vir_conv1.d3_conv1[0].in_channels)
            f, iset = 0, 0
            for i, (blk, s) in enumerate(zip((self.vir_conv1, self.vir_conv2, self.vir_conv3, self.vir_conv4), (1, 2, 4, 8))):
                f, iset = plan_nrconv(plan, blk, f, iset, s)
                plan.publish('x_conv%d' % (i + 1), f, iset)
            self._plan_mm_cache = plan
        return self._plan_mm_cache

    @staticmethod
    def _plan_usable(plan, feats):
        if (not executor.ENABLED or not feats.is_cuda or feats.shape[0] == 0 or feats.requires_grad or not plan.is_current()
                or not plan.eligible()):
            return False
        mode = plan.layers[0][1].training
        return all(bn.training == mode for _, bn in plan.layers)

    def _run(self, plan, feats, coords, shape, batch_size, proj, names):
        ci = spconv._as_i32(coords)
        run, res = executor.run_plan(plan, feats, ci, shape, batch_size, proj, plan.layers[0][1].training,
                                     plan.layers[0][0].precision)
        idict = executor.LazyIndiceDict(run, ci, shape)
        return tuple(_published_tensor(res, k, batch_size, idict) for k in names)

    def _lidar_stream(self, feats, coords, shape, batch_size):
        plan = self._plan_lidar()
        if not plan.is_current():          # a sub-module was replaced after the plan was built
            self._plan_lidar_cache = None
            plan = self._plan_lidar()
        if self._plan_usable(plan, feats):
            return self._run(plan, feats, coords, shape, batch_size, None, ('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4', 'out'))
        x = spconv.SparseConvTensor(feats, coords.int(), shape, batch_size)
        x1 = self.conv1(self.conv_input(x))
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        return x1, x2, x3, x4, self.conv_out(x4)

    def _maybe_discard(self, t, batch_dict, layer):
        if not self.training or self.discard_mode != 'paper' or self.layer_discard_rate == 0:
            return t
        given = batch_dict.get('stvd_keep_rows')
        keep = given[layer] if given is not None else stvd_keep_rows(t.features.shape[0], self.layer_discard_rate)
        return discard_rows(t, keep)

    def forward(self, batch_dict):
        rot_num = batch_dict['transform_param'].shape[1] if 'transform_param' in batch_dict else 1
        batch_size = batch_dict['batch_size']
        sfx = [''] + [str(i) for i in range(1, rot_num)]
        strides = {'x_conv1': 1, 'x_conv2': 2, 'x_conv3': 4, 'x_conv4': 8}
        ops.stats_arena(batch_dict['voxel_features'].device).reset()
        if self.training:
            for s in sfx:
                x1, x2, x3, x4, out = self._lidar_stream(batch_dict['voxel_features' + s], batch_dict['voxel_coords' + s],
                                                         self.sparse_shape, batch_size)
                batch_dict.update({
                    'encoded_spconv_tensor' + s: out, 'encoded_spconv_tensor_stride' + s: 8,
                    'multi_scale_3d_features' + s: {'x_conv1': x1, 'x_conv2': x2, 'x_conv3': x3, 'x_conv4': x4},
                    'multi_scale_3d_strides' + s: dict(strides)})
        else:
            feats, coords = [], []
            for i, s in enumerate(sfx):
                feats.append(batch_dict['voxel_features' + s])
                c = batch_dict['voxel_coords' + s].clone()
                c[:, 3] += i * self.sparse_shape[2]
                coords.append(c)
            big = [self.sparse_shape[0], self.sparse_shape[1], self.sparse_shape[2] * 4]      # factor 4 is hard-coded (:360)
            _, _, x3, x4, out = self._lidar_stream(torch.cat(feats, 0), torch.cat(coords), big, batch_size)
            for i, s in enumerate(sfx):
                batch_dict.update({
                    'encoded_spconv_tensor' + s: decompose_tensor(out, i, batch_size),
                    'encoded_spconv_tensor_stride' + s: 8,
                    'multi_scale_3d_features' + s: {'x_conv1': None, 'x_conv2': None,
                                                    'x_conv3': decompose_tensor(x3, i, batch_size),
                                                    'x_conv4': decompose_tensor(x4, i, batch_size)},
                    'multi_scale_3d_strides' + s: dict(strides)})
        if self.mm:
            calib = batch_dict['calib']
            for i, s in enumerate(sfx):
                feats, coords = batch_dict['voxel_features_mm' + s], batch_dict['voxel_coords_mm' + s]
                trans = batch_dict['aug_param'] if 'aug_param' in batch_dict else None
                if 'transform_param' in batch_dict:
                    trans = batch_dict['transform_param'][:, i, :]
                discarding = self.training and self.discard_mode == 'paper' and self.layer_discard_rate != 0
                if not self._plan_mm().is_current():
                    self._plan_mm_cache = None
                if not discarding and self._plan_usable(self._plan_mm(), feats):
                    side = ops.side(feats.device).stream if executor.TWO_STREAMS else None
                    proj = ops.projection_params(calib, trans, batch_size, feats.device, side)
                    x1, x2, x3, x4 = self._run(self._plan_mm(), feats, coords, self.sparse_shape, batch_size, proj,
                                               ('x_conv1', 'x_conv2', 'x_conv3', 'x_conv4'))
                    batch_dict.update({
                        'encoded_spconv_tensor_stride_mm' + s: 8,
                        'multi_scale_3d_features_mm' + s: {'x_conv1': x1, 'x_conv2': x2, 'x_conv3': x3, 'x_conv4': x4},
                        'multi_scale_3d_strides' + s: dict(strides)})
                    continue
                proj = ops.projection_params(calib, trans, batch_size, feats.device)
                x = spconv.SparseConvTensor(feats, coords.int(), self.sparse_shape, batch_size)
                x = self._maybe_discard(x, batch_dict, 0)              # VirConv8x also discards the input voxels (:489-490)
                x1 = self._maybe_discard(self.vir_conv1(x, batch_size, proj, 1), batch_dict, 1)
                x2 = self._maybe_discard(self.vir_conv2(x1, batch_size, proj, 2), batch_dict, 2)
                x3 = self._maybe_discard(self.vir_conv3(x2, batch_size, proj, 4), batch_dict, 3)
                x4 = self.vir_conv4(x3, batch_size, proj, 8)
                batch_dict.update({
                    'encoded_spconv_tensor_stride_mm' + s: 8,
                    'multi_scale_3d_features_mm' + s: {'x_conv1': x1, 'x_conv2': x2, 'x_conv3': x3, 'x_conv4': x4},
                    'multi_scale_3d_strides' + s: dict(strides)})
        return batch_dict
