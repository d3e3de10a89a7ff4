"""Plan executor front end: describe a backbone's layer graph once, run its whole forward / backward through ONE
C-ABI call each (`vc_exec_forward` / `vc_exec_backward`, include/virconv_b200.h; host loop in csrc/executor.cu).

The reference walks the backbone layer by layer in Python (`VirConvL8x.forward` spconv_backbone.py:609-699,
`NRConvBlock.forward` :207-229, `VirConv8x.forward` :339-535) and autograd walks it back; on a B200 that host work is
longer than the kernels it launches.  `Plan` records the same graph as a flat op list; `run_plan` executes it natively
and returns ordinary tensors (features with autograd history, indices) that live in one arena allocated per step.
"""
from __future__ import annotations

import ctypes
import os
import weakref

import numpy as np
import torch

from . import _lib, ops
from ._lib import check

ENABLED = os.environ.get('VIRCONV_EXECUTOR', '1') != '0'
TWO_STREAMS = os.environ.get('VIRCONV_EXEC_STREAMS', '3') != '1'
THREE_STREAMS = os.environ.get('VIRCONV_EXEC_STREAMS', '3') == '3'
# weight gradients on their own stream next to the BN-backward / dgrad chain (vc_exec_backward's wgrad_stream), with the
# wgrad CTAs limited to WGRAD_CTAS SMs they keep to themselves (vc_conv_wgrad_tc_config)
# (measured, profiles/sweep_wgrad_r1.txt: 3.91 ms/step without, 3.29-3.31 with 96-112 CTAs; the shared-memory floor that
# keeps gather CTAs off the wgrad SMs made no difference at 96 CTAs, so it is off)
WGRAD_STREAM = os.environ.get('VIRCONV_WGRAD_STREAM', '1') != '0'
WGRAD_CTAS = int(os.environ.get('VIRCONV_WGRAD_CTAS', '128'))     # re-swept after the dgrad chain got faster: sweep_wgrad_r1b.txt
WGRAD_SMEM_KB = int(os.environ.get('VIRCONV_WGRAD_SMEM_KB', '0'))
_WSTREAM = {}
_WCFG = [None]


def _wgrad_stream(device):
    key = (device.type, device.index)
    if key not in _WSTREAM:
        _WSTREAM[key] = torch.cuda.Stream(device=device)
    cfg = (WGRAD_CTAS, WGRAD_SMEM_KB * 1024) if WGRAD_STREAM else (148, 0)
    if _WCFG[0] != cfg:
        check(_lib.load().vc_conv_wgrad_tc_config(*cfg), 'vc_conv_wgrad_tc_config')
        _WCFG[0] = cfg
    return _WSTREAM[key] if WGRAD_STREAM else None


TIMING = False                     # bench.py's roofline pass: per-launch events around the conv kernels
LAST_RUN = None                    # the most recent forward's record while TIMING is on (bench.py reads its row counts)

OP_SUBM_RB, OP_CONV_RB, OP_INDEX2UV, OP_CBR, OP_CAT = 1, 2, 3, 4, 5
VC_ERR_WORKSPACE = -3
OPI, OPF, PCOLS = 24, 8, 9
KIND_NAMES = {0: 'conv_fwd', 1: 'conv_fwd_tc', 2: 'conv_dgrad', 3: 'conv_dgrad_tc', 4: 'conv_dgrad_scatter',
              5: 'conv_wgrad', 6: 'conv_wgrad_tc', 7: 'conv_dgrad_scatter_tc'}


def _t3(v, nd, fill):
    v = list(v) if isinstance(v, (list, tuple)) else [v] * nd
    assert len(v) == nd
    return [int(x) for x in v] + [fill] * (3 - nd)


class Plan:
    """Flat op list over numbered feature slots / index sets / rulebooks (slot 0 / set 0 = the network input)."""

    def __init__(self, in_channels):
        self.rows_i, self.rows_f = [], []
        self.layers = []              # (conv module, bn module)
        self.owners = []              # per layer: the SparseSequential the pair came from (or None)
        self.n_f, self.n_i, self.n_rb = 1, 1, 0
        self.f_channels = {0: in_channels}
        self.published = []           # (name, feature slot, index set)
        self.rb_keys = {}             # rulebook id -> (indice_keys, ndim, in index set, out index set)
        self._final = None
        # per-device run-time state (dies with the plan, i.e. with the model that owns it)
        self.arena_bytes = {}         # (device index, training) -> arena size handed out (monotone)
        self.arena_used = {}          # (device index, training) -> largest number of arena bytes a step really used
        self.reserved = {}            # (device, stream) -> arena size already parked in the allocator pool
        self.inflight = {}            # device index -> tail events of the forwards still in flight

    # a plan is a cache derived from the modules: copies / pickles of the owning model drop it and rebuild it lazily
    def __deepcopy__(self, memo):
        return None

    def __reduce__(self):
        return (type(None), ())

    def _row(self, kind, stream, a=0, b=0, c=0, ndim=0, ks=(0, 0, 0), st=(1, 1, 1), pd=(0, 0, 0), dl=(1, 1, 1), cin=0,
             cout=0, layer=0, x0=0, x1=0, x2=0, f=()):
        self.rows_i.append([kind, stream, a, b, c, ndim, *ks, *st, *pd, *dl, cin, cout, layer, x0, x1, x2])
        self.rows_f.append(list(f) + [0.0] * (OPF - len(f)))
        self._final = None

    def subm_rb(self, iset, ndim, ksize, dilation=1, unique=True, keys=(), stream=1):
        """stream: 1 = index stream, 2 = second index stream (the image branch: its projection + 2-D rulebooks depend on a
        stage's indices only, so they run beside the 3-D rulebooks of the same and the following stages)"""
        rb = self.n_rb
        self.n_rb += 1
        self.rb_keys[rb] = (tuple(keys), ndim, iset, iset)
        self._row(OP_SUBM_RB, stream, a=iset, c=rb, ndim=ndim, ks=_t3(ksize, ndim, 1), dl=_t3(dilation, ndim, 1), x0=int(unique))
        return rb

    def conv_rb(self, iset, ndim, ksize, stride, padding, dilation=1, keys=()):
        rb, out = self.n_rb, self.n_i
        self.n_rb += 1
        self.n_i += 1
        self.rb_keys[rb] = (tuple(keys), ndim, iset, out)
        self._row(OP_CONV_RB, 1, a=iset, b=out, c=rb, ndim=ndim, ks=_t3(ksize, ndim, 1), st=_t3(stride, ndim, 1),
                  pd=_t3(padding, ndim, 0), dl=_t3(dilation, ndim, 1))
        return out, rb

    def index2uv(self, iset, stride, pts_range=(0, -40, -3, 70.4, 40, 1), voxel_size=(0.05, 0.05, 0.05), u_max=1400,
                 v_max=600, image_shape=(1600, 600), stream=1):
        out = self.n_i
        self.n_i += 1
        vs = np.array(voxel_size, dtype=np.float64) * stride      # same host arithmetic as ops.index2uv
        grid = [vs[0], vs[1], vs[2], pts_range[0] + vs[0] / 2, pts_range[1] + vs[1] / 2, pts_range[2] + vs[2] / 2]
        self._row(OP_INDEX2UV, stream, a=iset, b=out, ndim=3, ks=(int(image_shape[0]), int(image_shape[1]), 0), x0=int(stride),
                  x1=int(u_max), x2=int(v_max), f=[float(np.float32(g)) for g in grid])
        return out

    def cbr(self, f_in, rb, conv, bn, seq=None):
        """`seq`: the container whose members [0], [1] the (conv, bn) pair was read from — lets is_current() notice a
        swapped sub-module (e.g. SyncBatchNorm.convert_sync_batchnorm after the plan was built, tools/train.py:115-116)."""
        out = self.n_f
        self.n_f += 1
        layer = len(self.layers)
        self.layers.append((conv, bn))
        self.owners.append(seq)
        assert self.f_channels[f_in] == conv.in_channels, (self.f_channels[f_in], conv.in_channels)
        self.f_channels[out] = conv.out_channels
        self._row(OP_CBR, 0, a=f_in, b=out, c=rb, cin=conv.in_channels, cout=conv.out_channels, layer=layer,
                  x0=int(f_in != 0))
        return out

    def cat(self, fa, fb):
        out = self.n_f
        self.n_f += 1
        self.f_channels[out] = self.f_channels[fa] + self.f_channels[fb]
        self._row(OP_CAT, 0, a=fa, b=fb, c=out)
        return out

    def publish(self, name, f_slot, iset):
        self.published.append((name, f_slot, iset))

    def finalize(self):
        if self._final is None:
            oi = np.ascontiguousarray(np.array(self.rows_i, dtype=np.int32).reshape(-1, OPI))
            of = np.ascontiguousarray(np.array(self.rows_f, dtype=np.float32).reshape(-1, OPF))
            lf = np.ascontiguousarray(np.array([[bn.eps, 0.0 if bn.momentum is None else bn.momentum] for _, bn in self.layers],
                                               dtype=np.float32).reshape(-1, 2))
            sizes = []
            for conv, bn in self.layers:
                sizes += [conv.weight.numel(), bn.weight.numel(), bn.bias.numel()]
            offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            self._final = (oi, of, lf, sizes, offs)
        return self._final

    def params(self):
        out = []
        for conv, bn in self.layers:
            out += [conv.weight, bn.weight, bn.bias]
        return out

    def eligible(self):
        """Every layer is conv(bias=False) + a plain affine BatchNorm1d with running statistics and a fixed momentum (what the
        VirConv blocks build), all parameters / buffers fp32 and contiguous.  SyncBatchNorm (statistics across ranks), a
        cumulative-average BatchNorm (momentum=None) or a half-precision model go through the module path instead."""
        def f32(*ts):
            return all(t is not None and t.dtype == torch.float32 and t.is_contiguous() for t in ts)
        return all(conv.bias is None and type(bn) is torch.nn.BatchNorm1d and bn.affine and bn.track_running_stats
                   and bn.momentum is not None and f32(conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)
                   for conv, bn in self.layers)

    def is_current(self):
        """The (conv, bn) objects recorded at build time are still the live members of their containers."""
        for (conv, bn), seq in zip(self.layers, self.owners):
            if seq is not None and (seq[0] is not conv or seq[1] is not bn):
                return False
        return True


_PINNED = {}


def _pinned(device):
    key = (device.type, device.index)
    if key not in _PINNED:
        _PINNED[key] = torch.empty(16, dtype=torch.int32).pin_memory()
    return _PINNED[key]


ARENA_GRAIN = 64 << 20


def _arena_bytes(plan, n0, device, with_bwd=True):
    """Arena size for a step.  First use of a plan: a generous guess from the input row count (VirConv-L uses ~16 KB per
    input voxel, forward + backward; a forward alone about half).  Afterwards: the largest bytes-per-input-row an earlier
    step of the same plan really used (recorded by the backward / the eval forward) times this step's rows plus 30 %
    head-room — monotone per plan, so the caching allocator keeps handing back the same block, and an eval-only or
    small-batch user does not reserve the training worst case (rows*24 KB + 256 MB over-reserved ~2x)."""
    key = (device.index, bool(with_bwd))
    per_row = plan.arena_used.get(key, 0.0)
    if per_row:
        want = int(1.3 * per_row * int(n0)) + (32 << 20)
    else:
        # (fixed part: bitmaps / hash tables / weight images / per-layer scratch do not shrink with the batch — a 14 k-voxel
        #  training step uses 540 MB, an 80 k-voxel one 1.1 GB)
        want = int(n0) * (24576 if with_bwd else 12288) + ((384 if with_bwd else 192) << 20)
    want = max(want, plan.arena_bytes.get(key, 0))
    want = (want + ARENA_GRAIN - 1) // ARENA_GRAIN * ARENA_GRAIN
    plan.arena_bytes[key] = want
    return want


def _note_arena_use(plan, device, with_bwd, used, n0):
    key = (device.index, bool(with_bwd))
    plan.arena_used[key] = max(plan.arena_used.get(key, 0.0), float(used) / max(int(n0), 1))


MAX_STEPS_AHEAD = 4          # the host never enqueues more than this many forwards of a plan beyond the GPU
ARENAS_IN_FLIGHT = MAX_STEPS_AHEAD + 3   # (+1 for the step being enqueued, +2: a block handed to three streams with record_stream
                                         #  returns to the pool only after all of them passed it — one cudaMalloc of a 2.3 GB arena in a
                                         #  timed loop costs 46 ms, profiles/bench_r2_eager_bf16.json vs its e2e outlier run)


def _throttle(plan, dev):
    """Bound how far the host runs ahead of the GPU.  Without a per-step synchronisation the only brake is the driver's
    launch-queue depth, and with the step's work spread over four streams that allowed more steps in flight than
    ARENAS_IN_FLIGHT — the next arena then cost a multi-GB cudaMalloc (a 30-100 ms stall once every few runs).  Each
    forward leaves an event at its tail; before a new forward starts, the host waits for the one MAX_STEPS_AHEAD steps
    back (free when the GPU is the bottleneck: the host would otherwise sit in a full launch queue anyway)."""
    q = plan.inflight.setdefault(dev.index, [])
    while len(q) >= MAX_STEPS_AHEAD:
        q.pop(0).synchronize()
    return q


def _alloc_arena(plan, nbytes, dev, stream):
    """One arena for this step.  The host may run several steps ahead of the GPU (no per-step synchronisation), so that
    many arenas are alive at once; a fresh multi-GB block costs a cudaMalloc of tens of milliseconds when it lands in
    the middle of training (seen as 5-10 ms/step outliers).  The first time a (plan, size) is used, ARENAS_IN_FLIGHT
    blocks are therefore allocated and released at once, which parks them in the caching allocator's pool of `stream`;
    every later request is served from there."""
    key = (dev.index, stream.cuda_stream if stream is not None else 0)
    ctx = torch.cuda.stream(stream) if stream is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        if plan.reserved.get(key) != nbytes:
            hold = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(ARENAS_IN_FLIGHT)]
            del hold
            plan.reserved[key] = nbytes
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)


_POOLED = {}


def reserve_blocks(tag, nbytes, dev, stream=None, count=16):
    """Park `count` blocks of `nbytes` in the caching allocator's pool of `stream` (once per tag/size).  For the few
    per-step tensors that are used on more than one stream (`record_stream` defers their reuse, so how many are alive at
    once depends on timing): without this the pool occasionally grows by a cudaMalloc in the middle of training, which
    stalls the busy GPU for tens of milliseconds."""
    key = (tag, dev.index, stream.cuda_stream if stream is not None else 0)
    if _POOLED.get(key, 0) >= nbytes:
        return
    ctx = torch.cuda.stream(stream) if stream is not None else None
    if ctx is not None:
        ctx.__enter__()
    try:
        hold = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(count)]
        del hold
    finally:
        if ctx is not None:
            ctx.__exit__(None, None, None)
    _POOLED[key] = nbytes


def _view(arena, ptr, shape, dtype):
    """A tensor over arena memory at device address `ptr` that is NOT an autograd view of the arena (Tensor.set_ on the
    shared storage), so it can be returned from an autograd Function and outlive the arena tensor object."""
    esz = torch.empty(0, dtype=dtype).element_size()
    off = ptr - arena.data_ptr()
    assert off >= 0 and off % esz == 0
    n = 1
    for s in shape:
        n *= int(s)
    assert off + n * esz <= arena.numel()
    t = torch.empty(0, dtype=dtype, device=arena.device)
    strides, acc = [], 1
    for s in reversed(shape):
        strides.append(acc)
        acc *= int(s)
    t.set_(arena.untyped_storage(), off // esz + arena.storage_offset() // esz, tuple(int(s) for s in shape),
           tuple(reversed(strides)))
    return t


class _Run:
    """One forward's record: arena + state blob, enough for the backward call and for queries."""

    def __init__(self, plan, arena, state, precision):
        self.plan, self.arena, self.state, self.precision = plan, arena, state, precision

    def query(self, what, idx):
        out = (ctypes.c_longlong * 8)()
        check(_lib.load().vc_exec_query(self.state.ctypes.data, what, idx, out), 'vc_exec_query')
        return [int(v) for v in out]

    def feature(self, slot):
        p, _, rows, c = self.query(1, slot)[:4]
        return _view(self.arena, p, (rows, c), torch.float32)

    def feature_bf16(self, slot):
        _, pb, rows, c = self.query(1, slot)[:4]
        return _view(self.arena, pb, (rows, c), torch.bfloat16) if pb else None

    def indices(self, iset):
        p, n, ndim, s0, s1, s2 = self.query(2, iset)[:6]
        return _view(self.arena, p, (n, 1 + ndim), torch.int32), [s0, s1, s2][:ndim]

    def count(self, iset):
        """Static mode: the device int32[1] row count of an index set (a view into the arena)."""
        p = self.query(2, iset)[6]
        return _view(self.arena, p, (1,), torch.int32) if p else None

    def rulebook(self, rb):
        """(nbr [K, n_out] int32, nbr_bwd or None, pair_num or None, meta dict) — tests compare these with the oracle."""
        p, pbw, ppn, K, n_in, n_out, subm, unique = self.query(3, rb)
        nbr = _view(self.arena, p, (K, n_out), torch.int32)
        nbw = _view(self.arena, pbw, (K, n_in), torch.int32) if pbw else None
        pn = _view(self.arena, ppn, (K,), torch.int32) if ppn else None
        return nbr, nbw, pn, dict(K=K, n_in=n_in, n_out=n_out, subm=bool(subm), unique=bool(unique))


class LazyIndiceDict(dict):
    """`SparseConvTensor.indice_dict` of the tensors a plan publishes: {indice_key: ops.Rulebook} like the module path
    leaves behind (spconv caches its indice pairs there), but materialised on first access — building a dozen tensor
    views per step for a dictionary nobody reads would be pure host overhead."""

    def __init__(self, run, coords0, shape0):
        super().__init__()
        self._src = (run, coords0, shape0)

    def _fill(self):
        if self._src is None:
            return
        run, coords0, shape0 = self._src
        self._src = None
        for rb, (keys, ndim, i_in, i_out) in run.plan.rb_keys.items():
            if not keys:
                continue
            nbr, nbw, pn, meta = run.rulebook(rb)
            oidx, oshape = (coords0, list(shape0)) if i_out == 0 else run.indices(i_out)
            book = ops.Rulebook(meta['subm'], ndim, meta['K'], meta['n_in'], meta['n_out'], nbr, nbw, pn, oidx, oshape,
                                meta['unique'])
            book._keepalive = run
            for k in keys:
                dict.__setitem__(self, k, book)

    def __getitem__(self, k):
        self._fill()
        return dict.__getitem__(self, k)

    def __contains__(self, k):
        self._fill()
        return dict.__contains__(self, k)

    def __iter__(self):
        self._fill()
        return dict.__iter__(self)

    def __len__(self):
        self._fill()
        return dict.__len__(self)

    def get(self, k, default=None):
        self._fill()
        return dict.get(self, k, default)

    def items(self):
        self._fill()
        return dict.items(self)

    def keys(self):
        self._fill()
        return dict.keys(self)

    def values(self):
        self._fill()
        return dict.values(self)


def _layer_ptrs(plan, grad_base=0):
    """[n_layers, 9] uint64 pointer table of vc_exec_*.  The parameter / buffer columns are cached per plan and re-read
    only when a weight's address changes (module.to(), load of a new tensor object); the gradient columns are
    `grad_base` + fixed offsets into the step's flat gradient buffer."""
    oi, of, lf, sizes, offs = plan.finalize()
    key = tuple((conv.weight.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                 bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr() if bn.num_batches_tracked is not None else 0)
                for conv, bn in plan.layers)
    cached = getattr(plan, '_ptr_cache', None)
    if cached is None or cached[0] != key:
        tab = np.zeros((len(plan.layers), PCOLS), dtype=np.uint64)
        for i, (conv, bn) in enumerate(plan.layers):
            nbt = bn.num_batches_tracked
            tab[i, 0] = conv.weight.data_ptr()
            tab[i, 1] = bn.weight.data_ptr()
            tab[i, 2] = bn.bias.data_ptr()
            tab[i, 3] = bn.running_mean.data_ptr()
            tab[i, 4] = bn.running_var.data_ptr()
            tab[i, 5] = nbt.data_ptr() if nbt is not None else 0
        goff = (4 * offs[:-1].reshape(-1, 3)).astype(np.uint64)
        plan._ptr_cache = cached = (key, tab, goff)
    tab = cached[1].copy()
    if grad_base:
        tab[:, 6:9] = cached[2] + np.uint64(grad_base)
    return tab


class StaticSpec:
    """Static (CUDA-graph capturable) execution of a plan: `n_dev` int32[1] device tensor = number of valid input rows
    (the feature / coordinate tensors passed to the plan are then CAPACITY sized), `caps` {index-set id: row capacity} for
    the strided convs' output sets, `overflow` int32[1] device tensor receiving the largest row count that did not fit."""

    def __init__(self, n_dev, caps, overflow, alias_params=False):
        self.n_dev, self.caps, self.overflow = n_dev, dict(caps), overflow
        self.alias_params = alias_params      # differentiate w.r.t. per-call leaf aliases of the parameters (run_plan)
        self.param_aliases = {}
        # phased execution (graph.PipelinedStep): 1 = only the index operators (run_index), 2 = only the feature operators
        # (run_plan), both in the caller-owned `arena` (deterministic bump allocation: the two phases agree on every address)
        self.phase, self.arena = 0, None
        self._arr = None

    def caps_array(self):
        if self._arr is None:
            a = np.zeros(64, dtype=np.int32)
            for k, v in self.caps.items():
                a[int(k)] = int(v)
            self._arr = a
        return self._arr


class PlanFn(torch.autograd.Function):
    """forward: (features of the published slots..., ) ; backward: gradients of every layer's weight / gamma / beta."""

    @staticmethod
    def forward(ctx, plan, holder, feats, coords, spatial_shape, batch_size, proj, training, precision, inputs_ready, static,
                with_bwd, *params):
        lib = _lib.load()
        ctx.set_materialize_grads(False)     # published tensors the loss does not touch arrive as None, not as zeros
        dev = feats.device
        oi, of, lf, sizes, offs = plan.finalize()
        feats = feats.contiguous()
        n0 = feats.shape[0]
        main = ops._stream()
        side_obj = ops.side(dev).stream if TWO_STREAMS else None
        side = side_obj.cuda_stream if side_obj is not None else None
        # second index stream (image-branch projection + 2-D rulebooks).  Only where the call's own fork / join orders it
        # (static mode, or inputs the side streams must wait for anyway): with `inputs_ready` the index streams do not wait
        # for main, and the arena / coordinate tensors are only registered with ONE side stream
        side2_obj = ops.side2(dev) if (side_obj is not None and THREE_STREAMS and not inputs_ready) else None
        side2 = side2_obj.cuda_stream if side2_obj is not None else None
        inflight = None
        if static is not None:
            inputs_ready = False             # under capture the side stream must fork from the capturing stream
        else:
            inflight = _throttle(plan, dev)
        tab = _layer_ptrs(plan)
        for attempt in range(3):
            nbytes = _arena_bytes(plan, n0, dev, with_bwd)
            if static is not None and static.arena is not None:
                arena = static.arena              # phased execution: the index phase already built its part in there
            elif static is not None:
                # (inside a capture: the graph's own pool.  No record_stream: the side / wgrad streams are forked from and
                #  joined back into the calling stream inside vc_exec_forward / vc_exec_backward, so every later use of the
                #  block in the calling stream's order comes after all of this step's uses)
                arena = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            elif side_obj is not None:
                # The arena belongs to the SIDE stream (its index kernels are the first writers, possibly while main still
                # runs the previous step); main's uses are registered with record_stream, so the caching allocator recycles
                # the block only after both streams are done with it.
                arena = _alloc_arena(plan, nbytes, dev, side_obj)
                arena.record_stream(torch.cuda.current_stream(dev))
            else:
                arena = _alloc_arena(plan, nbytes, dev, None)
            state = np.zeros(lib.vc_exec_state_bytes(), dtype=np.uint8)
            rc = lib.vc_exec_forward(oi.ctypes.data, of.ctypes.data, oi.shape[0], tab.ctypes.data, lf.ctypes.data, len(plan.layers),
                                     feats.data_ptr(), feats.shape[1], coords.data_ptr(), n0, _lib.host_i32(spatial_shape),
                                     int(batch_size), proj.data_ptr() if proj is not None else None, int(training),
                                     int(precision == 'bf16'), 1 | ((static.phase if static is not None else 0) << 8), arena.data_ptr(),
                                     arena.numel(),
                                     _pinned(dev).data_ptr(), ops.tc_error_flag(dev).data_ptr(), state.ctypes.data, state.size,
                                     main, side, 0 if (inputs_ready and side is not None) else 1,
                                     static.caps_array().ctypes.data if static is not None else None,
                                     static.n_dev.data_ptr() if static is not None else None,
                                     static.overflow.data_ptr() if static is not None else None, side2)
            if rc != VC_ERR_WORKSPACE:
                break
            # too small (first use of a plan on an unusually dense batch): the call is restartable — nothing it enqueued is
            # read by anyone — so take a bigger arena and run it again
            plan.arena_bytes[(dev.index, bool(with_bwd))] = 2 * arena.numel()
        check(rc, 'vc_exec_forward')
        if inflight is not None:
            tail = torch.cuda.Event()
            tail.record()
            inflight.append(tail)
        run = _Run(plan, arena, state, precision)
        run.static, run.n0, run.training, run.with_bwd = static, n0, bool(training), bool(with_bwd)
        holder.append(run)
        if not with_bwd:
            _note_arena_use(plan, dev, False, run.query(0, 0)[0], n0)
        if TIMING:
            global LAST_RUN
            LAST_RUN = run
        ctx.run, ctx.feats, ctx.coords, ctx.proj = run, feats, coords, proj
        return tuple(run.feature(slot) for _, slot, _ in plan.published)

    @staticmethod
    def backward(ctx, *grads):
        lib = _lib.load()
        run = ctx.run
        plan = run.plan
        oi, of, lf, sizes, offs = plan.finalize()
        dev = run.arena.device
        if run.static is None:
            reserve_blocks(('flat_grad', id(plan)), 4 * int(offs[-1]), dev)
        flat = torch.empty(int(offs[-1]), dtype=torch.float32, device=dev)
        tab = _layer_ptrs(plan, flat.data_ptr())
        gs = [None if g is None else g.contiguous() for g in grads]
        slots = np.array([slot for _, slot, _ in plan.published], dtype=np.int32)
        ext = np.array([0 if g is None else g.data_ptr() for g in gs], dtype=np.uint64)
        ws = _wgrad_stream(dev)
        if ws is not None and run.static is None:
            flat.record_stream(ws)
            run.arena.record_stream(ws)
        rc = lib.vc_exec_backward(oi.ctypes.data, of.ctypes.data, oi.shape[0], tab.ctypes.data, lf.ctypes.data, len(plan.layers),
                                  slots.ctypes.data, ext.ctypes.data, len(gs), run.arena.data_ptr(), run.arena.numel(),
                                  ops.tc_error_flag(dev).data_ptr(), run.state.ctypes.data, ops._stream(),
                                  ws.cuda_stream if ws is not None else None)
        if rc == VC_ERR_WORKSPACE:
            # (a forward taken without autograd's knowledge of a later backward: the next step's arena is sized for it)
            plan.arena_bytes[(dev.index, run.with_bwd)] = 2 * run.arena.numel()
        check(rc, 'vc_exec_backward')
        _note_arena_use(plan, dev, run.with_bwd, run.query(0, 0)[0], run.n0)
        run.flat_grad = flat
        views = flat.split(sizes)
        out = [v.view_as(p) for v, p in zip(views, plan.params())]
        return (None,) * 12 + tuple(out)


def run_index(plan, feats, coords_i32, spatial_shape, batch_size, proj, training, precision, static):
    """Phase 1 of a phased static execution: enqueue ONLY the index operators of the plan (rulebooks, projection) into
    `static.arena`; `run_plan` with `static.phase == 2` later enqueues the feature operators over the same arena."""
    ops._require_cuda(feats, coords_i32)
    assert static is not None and static.arena is not None and static.phase == 1
    lib = _lib.load()
    dev = feats.device
    oi, of, lf, sizes, offs = plan.finalize()
    tab = _layer_ptrs(plan)
    side_obj = ops.side(dev).stream if TWO_STREAMS else None
    side2_obj = ops.side2(dev) if (side_obj is not None and THREE_STREAMS) else None
    state = np.zeros(lib.vc_exec_state_bytes(), dtype=np.uint8)
    rc = lib.vc_exec_forward(oi.ctypes.data, of.ctypes.data, oi.shape[0], tab.ctypes.data, lf.ctypes.data, len(plan.layers),
                             feats.data_ptr(), feats.shape[1], coords_i32.data_ptr(), feats.shape[0], _lib.host_i32(spatial_shape),
                             int(batch_size), proj.data_ptr(), int(training), int(precision == 'bf16'), 1 | (1 << 8),
                             static.arena.data_ptr(), static.arena.numel(), _pinned(dev).data_ptr(), ops.tc_error_flag(dev).data_ptr(),
                             state.ctypes.data, state.size, ops._stream(), side_obj.cuda_stream if side_obj is not None else None, 1,
                             static.caps_array().ctypes.data, static.n_dev.data_ptr(), static.overflow.data_ptr(),
                             side2_obj.cuda_stream if side2_obj is not None else None)
    check(rc, 'vc_exec_forward (index phase)')


def run_plan(plan, feats, coords_i32, spatial_shape, batch_size, proj, training, precision, inputs_ready=False, static=None):
    """-> (run record, {name: (features, indices, spatial_shape, bf16 shadow, device row count or None)}).  `coords_i32`
    [N,4] int32 contiguous (b,z,y,x).  inputs_ready: the caller vouches that `coords_i32` and `proj` are complete in the
    SIDE stream's order (resident from an earlier step, or produced on / synchronised with ops.side(device).stream): the
    index pipeline then does not wait for the main stream and overlaps the previous step's backward.  static: a StaticSpec
    (capacity-sized tensors, device row counts, no host synchronisation: CUDA-graph capturable)."""
    ops._require_cuda(feats, coords_i32)
    assert coords_i32.dtype == torch.int32 and coords_i32.is_contiguous()
    holder = []
    # the arena also holds what a backward saves / stages: sized for it whenever autograd will record this call
    with_bwd = bool(training) or (torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in plan.params())))
    params = plan.params()
    if static is not None and with_bwd and static.alias_params:
        # Fresh leaf aliases of the parameters (same storage) for this call.  autograd keeps ONE AccumulateGrad node per leaf
        # alive as long as any earlier graph is, and that node remembers the stream it was created on: a node born in an eager
        # step on the legacy default stream makes the engine synchronise that stream with the capturing one — illegal during
        # CUDA-graph capture (cudaErrorStreamCaptureImplicit).  graph.GraphedStep differentiates w.r.t. the aliases.
        aliases = [p.detach().requires_grad_(True) for p in params]
        static.param_aliases = {id(p): a for p, a in zip(params, aliases)}
        params = aliases
    outs = PlanFn.apply(plan, holder, feats, coords_i32, list(spatial_shape), batch_size, proj, training, precision,
                        bool(inputs_ready), static, with_bwd, *params)
    run = holder[0]
    res = {}
    for (name, slot, iset), f in zip(plan.published, outs):
        idx, shape = (coords_i32, list(spatial_shape)) if iset == 0 else run.indices(iset)
        cnt = None
        if static is not None:
            cnt = static.n_dev if iset == 0 else run.count(iset)
        res[name] = (f, idx, shape, (lambda slot=slot: run.feature_bf16(slot)), cnt)    # shadow resolved on first use
    return run, res


class RunCounts:
    """What outlives a run record (whose arena goes back to the allocator with the step's tensors): the row count of every
    strided conv's output set, for sizing a StaticSpec."""

    def __init__(self, run):
        self.rows = {}
        if run.static is None:               # (a static run's counts live on the device)
            for rb, (keys, ndim, i_in, i_out) in run.plan.rb_keys.items():
                if i_out != i_in and i_out != 0:
                    self.rows[i_out] = int(run.query(2, i_out)[1])


_LAST_RUNS = weakref.WeakKeyDictionary()     # model -> RunCounts of its most recent exact-mode run


def note_last_run(model, run):
    if run.static is None:
        _LAST_RUNS[model] = RunCounts(run)


def last_run(model):
    return _LAST_RUNS.get(model)


def measured_caps(run, margin=1.3, grain=1024):
    """{index-set id: capacity} for a StaticSpec from an EXACT-mode run of the same plan on a representative batch (a run
    record or the RunCounts `last_run(model)` keeps): the observed row count of every strided conv's output set times
    `margin`, rounded up to `grain` rows."""
    rows = run.rows if isinstance(run, RunCounts) else RunCounts(run).rows
    return {i: (int(n * margin) + grain - 1) // grain * grain for i, n in rows.items()}


def timing_start():
    global TIMING
    TIMING = True
    check(_lib.load().vc_exec_timing(1), 'vc_exec_timing')


def timing_stop():
    """-> list of (kind name, layer, ms)."""
    global TIMING
    lib = _lib.load()
    n_max = 4096
    ms = (ctypes.c_float * n_max)()
    kl = (ctypes.c_int32 * (2 * n_max))()
    n = lib.vc_exec_timing_read(ms, kl, n_max)
    if n < 0:
        check(n, 'vc_exec_timing_read')
    TIMING = False
    check(lib.vc_exec_timing(0), 'vc_exec_timing')
    return [(KIND_NAMES[kl[2 * i]], int(kl[2 * i + 1]), float(ms[i])) for i in range(n)]


def alg_bytes_flops(run, kind, layer):
    """Algorithmic bytes / FLOPs of one timed conv launch (same formulas as ops.py's per-call accounting):
    every input row read once, every output row written once, one (in, out) pair per rulebook entry, the weights."""
    rb_id, in_slot, out_slot = run.query(4, layer)[4:7]
    _, _, pn, meta = run.rulebook(rb_id)
    conv = run.plan.layers[layer][0]
    cin, cout, K = conv.in_channels, conv.out_channels, meta['K']
    n_in, n_out, P = meta['n_in'], meta['n_out'], int(pn.sum().item())
    w = K * cin * cout
    if kind == 'conv_fwd_tc':
        b = n_in * cin * 2 + n_out * cout * 4 + w * 2
    elif kind in ('conv_dgrad_tc', 'conv_dgrad_scatter_tc'):
        b = n_out * cout * 2 + n_in * cin * 4 + w * 2
    elif kind == 'conv_wgrad_tc':
        b = (n_in * cin + n_out * cout) * 2 + w * 4
    else:
        b = (n_in * cin + n_out * cout + w) * 4
    return b + P * 8, 2 * P * cin * cout
