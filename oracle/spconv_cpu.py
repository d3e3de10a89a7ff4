"""A CPU `spconv.pytorch` work-alike built on the oracle rulebooks.  TEST INFRASTRUCTURE
(see oracle/__init__.py) — also the CPU baseline timed by `bench.py --impl reference`.

It restates spconv 2.1's *Native* algorithm ("CPU indexing + per-offset GEMM"):
rulebook on the CPU, then for every kernel offset gather rows -> `torch.mm` -> `index_add_`;
the submanifold centre offset is one dense mm over all rows.  Backward comes from torch autograd
over those same index ops (dgrad = per-offset mm with W^T scattered to inputs, wgrad =
gathered_in^T @ gathered_dout), so it is the reference algorithm's gradient, not a second
hand-written one.

API surface mirrored (what the reference imports through `pcdet/utils/spconv_utils.py:33-36` and
uses in `spconv_backbone.py`): SparseConvTensor(features, indices, spatial_shape, batch_size),
.replace_feature, .dense, SubMConv{2,3}d, SparseConv{2,3}d, SparseSequential, SparseModule,
conv.SparseConvolution.  Weight layout is spconv 2.x's (C_out, *kernel, C_in)
(`detector3d_template.py:358-370`).
"""
from __future__ import annotations

import math
import types

import numpy as np
import torch
import torch.nn as nn

from . import rulebook as rb

# 'numpy' = oracle/rulebook.py (the canonical definition); 'c' = oracle/csrc/rulebook_ref.c (sequential
# hash map like spconv's CPU indexing; identical tables, used for the timed CPU baseline)
RULEBOOK_BACKEND = 'numpy'


def _rulebook_impl():
    if RULEBOOK_BACKEND == 'c':
        from . import crulebook
        return crulebook
    return rb


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, indice_dict=None):
        self._features = features
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid

    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, val):
        # spconv 2.x forbids this (pcdet/utils/spconv_utils.py:58-64 relies on it)
        raise ValueError("you can't set feature directly, use 'x = x.replace_feature(your_new_feature)'")

    def replace_feature(self, feature):
        new = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid,
                               self.indice_dict)
        return new

    def dense(self, channels_first=True):
        idx = self.indices.long()
        shape = [self.batch_size] + list(self.spatial_shape) + [self._features.shape[1]]
        out = torch.zeros(shape, dtype=self._features.dtype, device=self._features.device)
        out[tuple(idx[:, i] for i in range(idx.shape[1]))] = self._features
        if not channels_first:
            return out
        nd = len(self.spatial_shape)
        return out.permute(0, nd + 1, *range(1, nd + 1)).contiguous()


class SparseModule(nn.Module):
    pass


def native_conv(features, weight, nbr, n_out, subm):
    """out[o] = sum_k in[nbr[k,o]] @ W[k];  weight (C_out, K, C_in)."""
    K = nbr.shape[0]
    centre = K // 2
    if subm:
        out = features @ weight[:, centre, :].t()
    else:
        out = features.new_zeros((n_out, weight.shape[0]))
    for k in range(K):
        if subm and k == centre:
            continue
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        if o.numel() == 0:
            continue
        i = nbr[k].index_select(0, o)
        out.index_add_(0, o, features.index_select(0, i) @ weight[:, k, :].t())   # scatter-add in place, like spconv Native
    return out


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, indice_key=None):
        super().__init__()
        assert groups == 1
        tup = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v,) * ndim
        self.ndim, self.in_channels, self.out_channels = ndim, in_channels, out_channels
        self.kernel_size, self.stride = tup(kernel_size), tup(stride)
        self.padding, self.dilation = tup(padding), tup(dilation)
        self.subm, self.indice_key = subm, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x: SparseConvTensor):
        key = self.indice_key
        cached = x.indice_dict.get(key) if key is not None else None
        idx_np = x.indices.detach().cpu().numpy()
        if self.subm:
            if cached is None:
                nbr = _rulebook_impl().subm_rulebook(idx_np, x.spatial_shape, self.kernel_size, self.dilation)
                cached = dict(nbr=torch.from_numpy(nbr).long(), nbr_np=nbr, out_indices=x.indices,
                              out_shape=x.spatial_shape)
                if key is not None:
                    x.indice_dict[key] = cached
            out_indices, out_shape = x.indices, x.spatial_shape
        else:
            if cached is None:
                oi, osh, nf, nb = _rulebook_impl().conv_rulebook(idx_np, x.spatial_shape, self.kernel_size, self.stride,
                                                   self.padding, self.dilation)
                cached = dict(nbr=torch.from_numpy(nf).long(), nbr_np=nf, nbr_bwd_np=nb,
                              out_indices=torch.from_numpy(oi).to(x.indices.dtype), out_shape=osh)
                if key is not None:
                    x.indice_dict[key] = cached
            out_indices, out_shape = cached['out_indices'], cached['out_shape']
        w = self.weight.reshape(self.out_channels, -1, self.in_channels)
        out = native_conv(x.features, w, cached['nbr'], out_indices.shape[0], self.subm)
        if self.bias is not None:
            out = out + self.bias
        res = SparseConvTensor(out, out_indices, out_shape, x.batch_size, x.grid, x.indice_dict)
        return res


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         True, indice_key)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         True, indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         False, indice_key)


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, **kw):
        super().__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         False, indice_key)


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x


def as_spconv_namespace():
    """Module objects that satisfy `import spconv.pytorch as spconv` / `spconv.conv.SparseConvolution`."""
    pkg = types.ModuleType('spconv')
    pt = types.ModuleType('spconv.pytorch')
    conv = types.ModuleType('spconv.pytorch.conv')
    conv.SparseConvolution = SparseConvolution
    for name in ('SparseConvTensor', 'SparseModule', 'SparseSequential', 'SubMConv3d', 'SubMConv2d',
                 'SparseConv3d', 'SparseConv2d', 'SparseConvolution'):
        setattr(pt, name, globals()[name])
    pt.conv = conv
    pkg.pytorch = pt
    pkg.__path__ = []
    pt.__path__ = []
    return {'spconv': pkg, 'spconv.pytorch': pt, 'spconv.pytorch.conv': conv}
