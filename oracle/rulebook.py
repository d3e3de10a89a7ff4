"""Rulebook ("indice pairs") construction — CPU oracle (numpy).  TEST INFRASTRUCTURE, see oracle/__init__.py.

The arithmetic restated here is spconv 2.1.22's `ops.get_indice_pairs` (third-party, not in
/root/reference).  Reference call sites that fix the geometry: `spconv_backbone.py:89`
(SubMConv3d, k=3), `:92-93` (SparseConv3d stride 2, pad 1 | (0,1,1)), `:113` (SubMConv2d on
`[1600,600]` image indices, `:217-222`), `:563-564` (SparseConv3d (3,1,1)/(2,1,1)/0).

Canonical form — this file IS the definition the CUDA path is compared against:
  * kernel offsets are numbered z-major: k = (kz*Ky + ky)*Kx + kx;
  * cross-correlation, like torch.nn.functional.conv3d:
        coord_in = coord_out*stride - pad + k*dilation
  * the rulebook is the *neighbour table* nbr[k, o] = input row feeding output row o through
    offset k, or -1; the spconv-style pair list is its per-offset compaction ordered by
    ascending output row (`pairs_from_nbr`);
  * submanifold conv: output rows == input rows (same order).  The centre offset is the identity
    (spconv's Native path does `out = features @ W[centre]` on all rows); every other offset
    resolves a coordinate to the LOWEST row index holding it (matters only for the 2-D image
    branch, whose indices contain duplicates, `spconv_backbone.py:215-222`);
  * regular (strided) conv: output rows are all coordinates reached by >=1 active input, ordered
    by ascending linear index ((b*D+z)*H+y)*W+x — batch most significant, which is also the
    batch-contiguity contract of the RoI head (`voxel_query_utils.py:85-91`).
"""
from __future__ import annotations

import numpy as np


def _tuple(v, nd):
    if isinstance(v, (list, tuple, np.ndarray)):
        assert len(v) == nd
        return tuple(int(x) for x in v)
    return (int(v),) * nd


def kernel_offsets(ksize):
    """[K, nd] offsets, first spatial dim slowest."""
    grids = np.meshgrid(*[np.arange(k) for k in ksize], indexing='ij')
    return np.stack([g.ravel() for g in grids], axis=1).astype(np.int64)


def linearize(indices, spatial_shape):
    idx = np.asarray(indices, dtype=np.int64)
    lin = idx[:, 0]
    for d, s in enumerate(spatial_shape):
        lin = lin * int(s) + idx[:, 1 + d]
    return lin


def out_spatial_shape(spatial_shape, ksize, stride, padding, dilation):
    # floor((in + 2p - d*(k-1) - 1)/s) + 1 ; gives 81->41->21->10->4 for the VirConv-L chain
    return [(int(i) + 2 * p - d * (k - 1) - 1) // s + 1
            for i, k, s, p, d in zip(spatial_shape, ksize, stride, padding, dilation)]


def _lookup_table(indices, spatial_shape):
    """sorted unique linear keys + the lowest row holding each."""
    lin = linearize(indices, spatial_shape)
    order = np.argsort(lin, kind='stable')
    slin = lin[order]
    uniq, first = np.unique(slin, return_index=True)
    return uniq, order[first]


def subm_rulebook(indices, spatial_shape, ksize, dilation=1):
    """Neighbour table nbr[K, N] (int32) of a submanifold convolution."""
    idx = np.asarray(indices, dtype=np.int64)
    n, nd = idx.shape[0], idx.shape[1] - 1
    ksize, dilation = _tuple(ksize, nd), _tuple(dilation, nd)
    offs = kernel_offsets(ksize)
    centre = np.array([k // 2 for k in ksize], dtype=np.int64)
    shape = np.asarray(spatial_shape, dtype=np.int64)
    uniq, winner = _lookup_table(idx, spatial_shape)
    nbr = np.full((offs.shape[0], n), -1, dtype=np.int32)
    for k, off in enumerate(offs):
        if np.all(off == centre):
            nbr[k] = np.arange(n, dtype=np.int32)
            continue
        c = idx[:, 1:] + (off - centre) * np.asarray(dilation)
        ok = np.all((c >= 0) & (c < shape), axis=1)
        lin = linearize(np.concatenate([idx[:, :1], c], axis=1), spatial_shape)
        pos = np.searchsorted(uniq, lin)
        pos_c = np.minimum(pos, uniq.size - 1) if uniq.size else pos
        hit = ok & (pos < uniq.size)
        if uniq.size:
            hit &= uniq[pos_c] == lin
            nbr[k, hit] = winner[pos_c[hit]]
    return nbr


def conv_rulebook(indices, spatial_shape, ksize, stride=1, padding=0, dilation=1):
    """Regular sparse convolution.  Returns (out_indices [M,1+nd] i32, out_shape,
    nbr_fwd [K, M] (input row per output row), nbr_bwd [K, N] (output row per input row))."""
    idx = np.asarray(indices, dtype=np.int64)
    n, nd = idx.shape[0], idx.shape[1] - 1
    ksize, stride = _tuple(ksize, nd), _tuple(stride, nd)
    padding, dilation = _tuple(padding, nd), _tuple(dilation, nd)
    oshape = out_spatial_shape(spatial_shape, ksize, stride, padding, dilation)
    offs = kernel_offsets(ksize)
    K = offs.shape[0]
    s, p, d = np.asarray(stride), np.asarray(padding), np.asarray(dilation)
    osh = np.asarray(oshape, dtype=np.int64)
    out_lin = np.full((K, n), -1, dtype=np.int64)
    for k, off in enumerate(offs):
        num = idx[:, 1:] + p - off * d
        ok = np.all(num % s == 0, axis=1)
        o = num // s
        ok &= np.all((o >= 0) & (o < osh), axis=1)
        lin = linearize(np.concatenate([idx[:, :1], o], axis=1), oshape)
        out_lin[k, ok] = lin[ok]
    uniq = np.unique(out_lin[out_lin >= 0])
    m = uniq.size
    out_idx = np.zeros((m, 1 + nd), dtype=np.int32)
    rem = uniq.copy()
    for dd in range(nd - 1, -1, -1):
        out_idx[:, 1 + dd] = rem % oshape[dd]
        rem //= oshape[dd]
    out_idx[:, 0] = rem
    nbr_bwd = np.full((K, n), -1, dtype=np.int32)
    nbr_fwd = np.full((K, m), -1, dtype=np.int32)
    rows = np.arange(n, dtype=np.int32)
    for k in range(K):
        ok = out_lin[k] >= 0
        orow = np.searchsorted(uniq, out_lin[k, ok]).astype(np.int32)
        nbr_bwd[k, ok] = orow
        nbr_fwd[k, orow] = rows[ok]
    return out_idx, oshape, nbr_fwd, nbr_bwd


def pairs_from_nbr(nbr):
    """spconv-style `indice_pairs [2, K, N]` (-1 padded) and `indice_pair_num [K]`, each offset's
    pairs ordered by ascending output row."""
    K, n = nbr.shape
    pairs = np.full((2, K, n), -1, dtype=np.int32)
    num = np.zeros(K, dtype=np.int32)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0].astype(np.int32)
        num[k] = o.size
        pairs[0, k, :o.size] = nbr[k, o]
        pairs[1, k, :o.size] = o
    return pairs, num


def subm_rulebook_sequential(indices, spatial_shape, ksize, dilation=1):
    """Pure-Python sequential hash-map statement of the same table (small cases only):
    first insertion wins, exactly the order a CPU `unordered_map::insert` loop gives."""
    idx = np.asarray(indices, dtype=np.int64)
    n, nd = idx.shape[0], idx.shape[1] - 1
    ksize, dilation = _tuple(ksize, nd), _tuple(dilation, nd)
    table = {}
    for r in range(n):
        table.setdefault(tuple(idx[r]), r)
    offs = kernel_offsets(ksize)
    centre = tuple(k // 2 for k in ksize)
    nbr = np.full((offs.shape[0], n), -1, dtype=np.int32)
    for k, off in enumerate(offs):
        for r in range(n):
            if tuple(off) == centre:
                nbr[k, r] = r
                continue
            c = [int(idx[r, 0])]
            ok = True
            for dd in range(nd):
                v = int(idx[r, 1 + dd]) + (int(off[dd]) - centre[dd]) * dilation[dd]
                ok &= 0 <= v < int(spatial_shape[dd])
                c.append(v)
            if ok:
                nbr[k, r] = table.get(tuple(c), -1)
    return nbr
