"""Deterministic parameter filling shared by the golden-vector script and the tests.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np
import torch


def fill_module(module: torch.nn.Module, seed: int = 666, running_stats: bool = True):
    """Fill every parameter / BN buffer of `module` from one numpy Generator, visiting names in sorted
    order, so the reference model, the oracle model and the CUDA model (same state_dict keys) get
    bit-identical values without storing a checkpoint."""
    rng = np.random.default_rng(seed)
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd.keys()):
            t = sd[name]
            if name.endswith('num_batches_tracked'):
                t.zero_()
            elif name.endswith('running_var'):
                v = rng.uniform(0.5, 1.5, size=tuple(t.shape)) if running_stats else np.ones(tuple(t.shape))
                t.copy_(torch.from_numpy(v.astype(np.float32)))
            elif name.endswith('running_mean'):
                v = rng.normal(0, 0.1, size=tuple(t.shape)) if running_stats else np.zeros(tuple(t.shape))
                t.copy_(torch.from_numpy(v.astype(np.float32)))
            elif t.dim() == 1 and name.endswith('.weight'):      # BN gamma
                t.copy_(torch.from_numpy(rng.uniform(0.5, 1.5, size=tuple(t.shape)).astype(np.float32)))
            elif t.dim() == 1:                                   # BN beta / bias
                t.copy_(torch.from_numpy(rng.normal(0, 0.1, size=tuple(t.shape)).astype(np.float32)))
            else:                                                # conv weight (C_out, *k, C_in)
                fan_in = int(np.prod(t.shape[1:]))
                v = rng.normal(0, (2.0 / fan_in) ** 0.5, size=tuple(t.shape))
                t.copy_(torch.from_numpy(v.astype(np.float32)))
    return module


def rel_err(got, want):
    """max|got-want| / max|want| — the tolerance form BASELINE.json's north_star states (1e-4)."""
    got = torch.as_tensor(got).double()
    want = torch.as_tensor(want).double()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))


def join_by_coords(idx_a, feat_a, idx_b, feat_b):
    """Align two sparse tensors by coordinate (row order may differ); returns (feat_a, feat_b[perm])."""
    a = np.asarray(idx_a).astype(np.int64)
    b = np.asarray(idx_b).astype(np.int64)
    assert a.shape == b.shape, (a.shape, b.shape)
    mx = np.maximum(a.max(0), b.max(0)) + 1
    ka = np.ravel_multi_index(a.T, mx)
    kb = np.ravel_multi_index(b.T, mx)
    oa, ob = np.argsort(ka, kind='stable'), np.argsort(kb, kind='stable')
    assert np.array_equal(ka[oa], kb[ob]), 'coordinate sets differ'
    perm = np.empty_like(oa)
    perm[oa] = ob
    return feat_a, torch.as_tensor(feat_b)[torch.from_numpy(perm)]
