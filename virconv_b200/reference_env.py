"""Make the reference's OWN Python (`pcdet/`, `tools/train.py`, `tools/test.py`) importable on this stack (Python 3.12, torch
2.11, no spconv / cumm / compiled `pcdet.ops` extensions) — SURVEY §8f-2.

    import virconv_b200.reference_env as renv
    renv.install(use_fused_backbones=True)      # before anything from `pcdet` is imported

What it registers (nothing is written into the reference tree, nothing from it is copied):
  * `spconv`, `spconv.pytorch`, `spconv.utils`, `cumm.tensorview`      -> virconv_b200.spconv_compat (the sm_100a kernels)
  * Python-3.12 compatibility: `collections.Iterable` & friends (tools/train_utils/optimization/fastai_optim.py:3),
    `numpy.int/float/bool` aliases the 2021-era code still uses
  * small pure-Python stand-ins when the real package is absent: `easydict.EasyDict`, `tensorboardX.SummaryWriter`
    (falls back to a no-op writer), `prefetch_generator.BackgroundGenerator`, `skimage.transform` (import only)
  * the six compiled extensions of `pcdet/ops` (setup.py:52-128) as Python modules:
      - `pointnet2_stack_cuda.voxel_query_wrapper / group_points_wrapper / group_points_grad_wrapper` run on this
        library's kernels (vc_voxel_query, vc_group_points, vc_group_points_grad: bit-exact against the reference's own
        CUDA, tests/test_gpu_roi_pool.py) — the voxel-RoI pooling of TEDMHead (ted_head.py:496-552);
      - `roiaware_pool3d_cuda.points_in_boxes_cpu` (dataset-side augmentation) is restated in torch;
      - everything else (rotated-box NMS / IoU, ball query, FPS, deformable conv, VOTR ops) is import-only: calling
        it raises NotImplementedError naming the reference source that would have to be built.
With `use_fused_backbones` the detector registry (`pcdet.models.backbones_3d.__all__`) hands out this package's
VirConvL8x / VirConv8x (plan executor, fused conv+BN+ReLU) instead of the reference classes over the compat layer.
"""
from __future__ import annotations

import collections
import collections.abc
import importlib
import sys
import types

import numpy as np
import torch


class EasyDict(dict):
    """Attribute-style dict, nested (what `pcdet/config.py` needs from the `easydict` package)."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(EasyDict(x) if isinstance(x, dict) and not isinstance(x, EasyDict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k, v in d.items():
            setattr(self, k, v)

    def pop(self, k, *a):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *a)


class _NullWriter:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:       # noqa: BLE001
        return True


def _unbuilt(ext, fn, src):
    def f(*a, **k):
        raise NotImplementedError(f'{ext}.{fn}: this compiled extension of the reference ({src}) is not on the VirConv sparse-conv '
                                  f'hot path and has no sm_100a build in virconv_b200 (SURVEY §2b / §8f-2)')
    f.__name__ = fn
    return f


def _stub_ext(modname, src, fns, **real):
    attrs = {fn: _unbuilt(modname.rsplit('.', 1)[-1], fn, src) for fn in fns}
    attrs.update(real)
    m = types.ModuleType(modname)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[modname] = m
    parent, leaf = modname.rsplit('.', 1)
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


# ---- pointnet2_stack: the voxel-RoI pooling primitives, on this library's kernels ---------------------------------------
def _voxel_query_wrapper(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords, point_indices, idx):
    """pointnet2_stack_cuda.voxel_query_wrapper (voxel_query_utils.py:30-33): writes idx [M, nsample] in place."""
    from . import _lib, ops
    lib = _lib.load()
    _lib.check(lib.vc_voxel_query(int(M), int(R1), int(R2), int(R3), int(nsample), float(radius), int(z_range), int(y_range),
                                  int(x_range), ops._p(new_xyz), ops._p(xyz), ops._p(new_coords), ops._p(point_indices),
                                  ops._p(idx), None, ops._stream()), 'vc_voxel_query')


def _group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    """pointnet2_stack_cuda.group_points_wrapper (pointnet2_utils.py:74-75): out [M, C, nsample]."""
    from . import _lib, ops
    lib = _lib.load()
    _lib.check(lib.vc_group_points(int(B), int(M), int(C), int(nsample), ops._p(features), ops._p(features_batch_cnt), ops._p(idx),
                                   ops._p(idx_batch_cnt), ops._p(out), ops._stream()), 'vc_group_points')


def _group_points_grad_wrapper(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    """pointnet2_stack_cuda.group_points_grad_wrapper (pointnet2_utils.py:95-96): grad_features [N, C] (zeroed by the caller)."""
    from . import _lib, ops
    lib = _lib.load()
    _lib.check(lib.vc_group_points_grad(int(B), int(M), int(C), int(N), int(nsample), ops._p(grad_out), ops._p(idx),
                                        ops._p(idx_batch_cnt), ops._p(features_batch_cnt), ops._p(grad_features), ops._stream()),
               'vc_group_points_grad')


def _points_in_boxes_cpu(boxes, points, point_indices):
    """roiaware_pool3d_cuda.points_in_boxes_cpu (roiaware_pool3d_utils.py:9-26; src/roiaware_pool3d.cpp:106-140): CPU tensors,
    boxes [N, 7] (x, y, z, dx, dy, dz, heading) with z at the box centre, points [M, 3]; point_indices [N, M] int32 set to 1
    where the point lies inside the box (|local z| <= dz/2, |local x| < dx/2, |local y| < dy/2 after rotating by -heading)."""
    b = boxes.float()
    p = points.float()
    dz = (p[None, :, 2] - b[:, None, 2]).abs()
    sx, sy = p[None, :, 0] - b[:, None, 0], p[None, :, 1] - b[:, None, 1]
    c, s = torch.cos(-b[:, 6])[:, None], torch.sin(-b[:, 6])[:, None]
    lx = sx * c - sy * s
    ly = sx * s + sy * c
    inside = (dz <= b[:, None, 5] / 2) & (lx.abs() < b[:, None, 3] / 2) & (ly.abs() < b[:, None, 4] / 2)
    point_indices.copy_(inside.to(point_indices.dtype))
    return 1


def install(use_fused_backbones=False, reference_root=None):
    """Register everything listed in the module docstring.  `reference_root`: put that checkout first on sys.path."""
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from . import spconv_compat
    spconv_compat.install_as_spconv()
    # Python 3.10+: the ABC aliases left `collections`; numpy 1.24+: the builtin aliases left `numpy`
    for name in ('Iterable', 'Mapping', 'MutableMapping', 'Sequence', 'Callable'):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    for name, typ in (('int', int), ('float', float), ('bool', bool), ('object', object)):
        if name not in np.__dict__:
            try:
                setattr(np, name, typ)
            except Exception:       # noqa: BLE001
                pass
    if _missing('easydict'):
        _module('easydict', EasyDict=EasyDict)
    if _missing('tensorboardX'):
        writer = _NullWriter
        try:
            from torch.utils.tensorboard import SummaryWriter as writer     # noqa: N813
        except Exception:       # noqa: BLE001
            pass
        _module('tensorboardX', SummaryWriter=writer)
    if _missing('prefetch_generator'):
        class BackgroundGenerator:
            def __init__(self, generator, max_prefetch=1):
                self._g = iter(generator)

            def __iter__(self):
                return self._g

            def __next__(self):
                return next(self._g)
        _module('prefetch_generator', BackgroundGenerator=BackgroundGenerator)
    if _missing('skimage'):
        sk = _module('skimage')
        sk.transform = _module('skimage.transform', resize=_unbuilt('skimage.transform', 'resize', 'scikit-image'))
        sk.io = _module('skimage.io', imread=_unbuilt('skimage.io', 'imread', 'scikit-image'))

    # the compiled extensions of pcdet/ops: module objects under the names the reference imports
    _stub_ext('pcdet.ops.iou3d_nms.iou3d_nms_cuda', 'pcdet/ops/iou3d_nms/src',
              ['boxes_overlap_bev_gpu', 'boxes_iou_bev_gpu', 'nms_gpu', 'nms_normal_gpu', 'boxes_iou_bev_cpu'])
    _stub_ext('pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda', 'pcdet/ops/roiaware_pool3d/src',
              ['forward', 'backward', 'points_in_boxes_gpu'], points_in_boxes_cpu=_points_in_boxes_cpu)
    _stub_ext('pcdet.ops.roipoint_pool3d.roipoint_pool3d_cuda', 'pcdet/ops/roipoint_pool3d/src', ['forward'])
    _stub_ext('pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda', 'pcdet/ops/pointnet2/pointnet2_batch/src',
              ['ball_query_wrapper', 'group_points_wrapper', 'group_points_grad_wrapper', 'gather_points_wrapper',
               'gather_points_grad_wrapper', 'furthest_point_sampling_wrapper', 'three_nn_wrapper', 'three_interpolate_wrapper',
               'three_interpolate_grad_wrapper'])
    _stub_ext('pcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda', 'pcdet/ops/pointnet2/pointnet2_stack/src',
              ['ball_query_wrapper', 'furthest_point_sampling_wrapper', 'three_nn_wrapper', 'three_interpolate_wrapper',
               'three_interpolate_grad_wrapper', 'stack_furthest_point_sampling_wrapper', 'query_stacked_local_neighbor_idxs_wrapper_stack',
               'query_three_nn_by_stacked_local_idxs_wrapper_stack', 'vector_pool_wrapper', 'vector_pool_grad_wrapper'],
              voxel_query_wrapper=_voxel_query_wrapper, group_points_wrapper=_group_points_wrapper,
              group_points_grad_wrapper=_group_points_grad_wrapper)
    _stub_ext('pcdet.ops.votr_ops.votr_ops_cuda', 'pcdet/ops/votr_ops/src', [])
    _stub_ext('pcdet.ops.dcn.deform_conv_cuda', 'pcdet/ops/dcn/src', [])
    if use_fused_backbones:
        b3d = importlib.import_module('pcdet.models.backbones_3d')
        from .backbone import VirConv8x, VirConvL8x
        b3d.__all__['VirConvL8x'] = VirConvL8x
        b3d.__all__['VirConv8x'] = VirConv8x
    return True
