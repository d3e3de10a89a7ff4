"""ctypes front end of oracle/csrc/rulebook_ref.c — same outputs as oracle/rulebook.py, ~50x faster.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_int, c_int32, c_int64, c_void_p

import numpy as np

from . import cbuild
from . import rulebook as rb

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(cbuild.build())
        lib.ref_subm_rulebook.restype = c_int
        lib.ref_subm_rulebook.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
        lib.ref_conv_rulebook_count.restype = c_int64
        lib.ref_conv_rulebook_count.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                POINTER(c_void_p)]
        lib.ref_conv_rulebook_fill.restype = c_int
        lib.ref_conv_rulebook_fill.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
        lib.ref_free.argtypes = [c_void_p]
        _lib = lib
    return _lib


def _i32(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


def subm_rulebook(indices, spatial_shape, ksize, dilation=1):
    lib = _load()
    idx = _i32(indices)
    n, nd = idx.shape[0], idx.shape[1] - 1
    ks, dil, shape = _i32(rb._tuple(ksize, nd)), _i32(rb._tuple(dilation, nd)), _i32(spatial_shape)
    nbr = np.empty((int(np.prod(ks)), n), dtype=np.int32)
    assert lib.ref_subm_rulebook(idx.ctypes.data, n, nd, shape.ctypes.data, ks.ctypes.data, dil.ctypes.data,
                                 nbr.ctypes.data) == 0
    return nbr


def conv_rulebook(indices, spatial_shape, ksize, stride=1, padding=0, dilation=1):
    lib = _load()
    idx = _i32(indices)
    n, nd = idx.shape[0], idx.shape[1] - 1
    ks, st = rb._tuple(ksize, nd), rb._tuple(stride, nd)
    pd, dil = rb._tuple(padding, nd), rb._tuple(dilation, nd)
    oshape = rb.out_spatial_shape(spatial_shape, ks, st, pd, dil)
    a = [_i32(oshape), _i32(ks), _i32(st), _i32(pd), _i32(dil)]
    ptrs = [x.ctypes.data for x in a]
    cells = c_void_p()
    m = lib.ref_conv_rulebook_count(idx.ctypes.data, n, nd, *ptrs, ctypes.byref(cells))
    assert m >= 0
    K = int(np.prod(ks))
    out_idx = np.empty((m, 1 + nd), dtype=np.int32)
    nbr_fwd = np.empty((K, m), dtype=np.int32)
    nbr_bwd = np.empty((K, n), dtype=np.int32)
    assert lib.ref_conv_rulebook_fill(idx.ctypes.data, n, nd, *ptrs, cells, m, out_idx.ctypes.data, nbr_fwd.ctypes.data,
                                      nbr_bwd.ctypes.data) == 0
    lib.ref_free(cells)
    return out_idx, oshape, nbr_fwd, nbr_bwd
