"""Whole training step (backbone forward + loss + backward) as ONE CUDA graph.

The reference drives every step from Python (`tools/train_utils/train_utils.py:27-57`: `model(batch)`, `loss.backward()`):
~200 kernel launches whose enqueue cost (3 ms of host time per step in round 1) exceeds the kernels' run time on a B200.
The plan executor's static mode (`executor.StaticSpec`, `vc_exec_forward(..., caps, n0_dev, overflow)`) keeps every
data-dependent row count in device memory and sizes every buffer from host-side capacities, so nothing in a step reads a
device value on the host; `GraphedStep` captures such a step once and replays it:

    step = GraphedStep(model, loss_fn, params)          # model: VirConvL8x in train / eval mode
    for batch in loader:
        loss = step(batch)                              # copies the batch into the graph's input buffers, replays
        optimizer.step()                                # parameter .grad tensors are rewritten in place by every replay

Published tensors inside the graph are capacity sized: `t.features` [cap, C] with a zero tail, `t.indices` [cap, 4] with a
-1 tail, `t.num_rows` a device int32[1]; `loss_fn(batch_dict)` must use `num_rows` where the reference uses `shape[0]`.
A batch whose row counts exceed the captured capacities is detected through the executor's overflow flag (read back
asynchronously, one step late) and handled by re-capturing with larger capacities and re-running that batch.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, executor, ops


def masked_mean(t):
    """Mean over the valid rows of a published tensor: static mode (zero tail, device row count) or exact mode."""
    if getattr(t, 'num_rows', None) is None:
        return t.features.mean()
    return t.features.sum() / (t.num_rows.to(torch.float32) * t.features.shape[1]).squeeze(0)


class GraphedStep:
    def __init__(self, model, loss_fn, params=None, margin=1.3, grain=1024, warmup=2, check_overflow=True, voxelizer=None):
        """voxelizer: None — batches carry `voxel_features` / `voxel_coords` (what the reference's dataloader + MeanVFE
        deliver); or a dict(point_cloud_range, voxel_size, max_points_per_voxel, max_voxels, vfe_model) — batches carry the
        collated raw `points` [N, 1+C] and the hash-grid voxelisation + VFE (vc_voxelize_mean) run INSIDE the graph."""
        self.model, self.loss_fn = model, loss_fn
        self.vox = dict(voxelizer) if voxelizer is not None else None
        self.params = list(model.parameters()) if params is None else list(params)
        self.dev = self.params[0].device
        self.margin, self.grain, self.warmup = margin, grain, warmup
        self.check_overflow = check_overflow
        self.graph = None
        self.cap0 = 0
        self.caps = {}
        self.recaptures = 0
        self.launches_per_replay = 0
        self._pending = []        # (event, pinned slot) of the replays whose overflow flag has not been looked at yet
        self._slot = 0

    # ------------------------------------------------------------------------------------------------
    def _round(self, n):
        return (int(n * self.margin) + self.grain - 1) // self.grain * self.grain

    def _exact_step(self, batch):
        """One ordinary (exact-shape, host-synchronised) step; also measures the row counts of this batch."""
        for p in self.params:
            p.grad = None
        bd = dict(batch)
        if self.vox is not None:
            v = self.vox
            f, c, _ = ops.voxelize_mean(batch['points'], int(batch['batch_size']), v['point_cloud_range'], v['voxel_size'],
                                        v['max_points_per_voxel'], v['max_voxels'], v['vfe_model'])
            bd.update(voxel_features=f, voxel_coords=c)
        out = self.model(bd)
        loss = self.loss_fn(out)
        loss.backward()
        return loss.detach(), executor.last_run(self.model)

    def _make_buffers(self, batch, run):
        for k, v in executor.measured_caps(run, self.margin, self.grain).items():
            self.caps[k] = max(self.caps.get(k, 0), v)
        B = int(batch['batch_size'])
        if self.vox is not None:
            pts = batch['points']
            dev = self.params[0].device
            self.cap_pts = max(getattr(self, 'cap_pts', 0), self._round(pts.shape[0]))
            # unused tail rows: last sample's batch index, coordinates far outside the range (dropped by the voxeliser)
            row = torch.zeros(pts.shape[1], dtype=torch.float32)
            row[0] = B - 1
            row[1:4] = -1.0e6
            self.sentinel = row.to(dev)
            self.pts = self.sentinel.repeat(self.cap_pts, 1).contiguous()
            self._prev_n = 0
            self.cap0 = B * int(self.vox['max_voxels'])
        else:
            dev = batch['voxel_features'].device
            vf, vc = batch['voxel_features'], batch['voxel_coords']
            self.cap0 = max(self.cap0, self._round(vf.shape[0]))
            self.vf = torch.zeros((self.cap0, vf.shape[1]), dtype=torch.float32, device=dev)
            self.vc = torch.full((self.cap0, vc.shape[1]), -1, dtype=vc.dtype, device=dev)
        self.dev, self.B = dev, B
        self.n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.proj = torch.zeros((B, 28), dtype=torch.float32, device=dev)
        # pinned staging for the small per-step host values: a ring of slots, each guarded by an event recorded after its
        # upload — the host runs several replays ahead of the GPU, and rewriting ONE pinned buffer before the previous
        # step's asynchronous copy has executed would hand that step the next step's row count / projection
        self.STAGE_SLOTS = 8
        self.proj_host = torch.zeros((self.STAGE_SLOTS, B, 28), dtype=torch.float32).pin_memory()
        self.n_host = torch.zeros((self.STAGE_SLOTS, 1), dtype=torch.int32).pin_memory()
        self._stage_ev = [None] * self.STAGE_SLOTS
        self._stage = 0
        self.ovf_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.err_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._pending, self._slot = [], 0
        self.spec = executor.StaticSpec(self.n_dev, self.caps, self.overflow, alias_params=True)

    def _load(self, batch):
        """Copy one batch into the graph's input buffers (asynchronous, current stream)."""
        if self.vox is not None:
            pts = batch['points']
            n = pts.shape[0]
            if n > self.cap_pts:
                return False
            self.pts[:n].copy_(pts, non_blocking=True)
            if self._prev_n > n:                       # rows of the previous (larger) batch still sit behind this one
                self.pts[n:self._prev_n] = self.sentinel
            self._prev_n = n
        else:
            vf, vc = batch['voxel_features'], batch['voxel_coords']
            n = vf.shape[0]
            if n > self.cap0:
                return False
        slot = self._stage
        self._stage = (slot + 1) % self.STAGE_SLOTS
        if self._stage_ev[slot] is not None:
            self._stage_ev[slot].synchronize()         # (8 steps back: has long completed unless the GPU is far behind)
        if self.vox is None:
            self.vf[:n].copy_(vf, non_blocking=True)
            self.vc[:n].copy_(vc, non_blocking=True)
            self.n_host[slot, 0] = n
            self.n_dev.copy_(self.n_host[slot], non_blocking=True)
        trans = batch.get('aug_param')
        self.proj_host[slot].copy_(torch.from_numpy(ops.projection_params_host(batch['calib'], trans, self.B)))
        self.proj.copy_(self.proj_host[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._stage_ev[slot] = ev
        return True

    def _static_batch(self, batch):
        bd = {k: v for k, v in batch.items() if k not in ('voxel_features', 'voxel_coords', 'points')}
        if self.vox is not None:
            v = self.vox
            f, c, num, n_dev = ops.voxelize_mean(self.pts, self.B, v['point_cloud_range'], v['voxel_size'], v['max_points_per_voxel'],
                                                 v['max_voxels'], v['vfe_model'], static=True)
            spec = executor.StaticSpec(n_dev, self.caps, self.overflow, alias_params=True)
            bd.update(voxel_features=f, voxel_coords=c, voxel_num_points=num, virconv_static=spec, virconv_proj=self.proj)
        else:
            spec = self.spec
            bd.update(voxel_features=self.vf, voxel_coords=self.vc, virconv_static=spec, virconv_proj=self.proj)
        self._cur_spec = spec
        return bd

    def _static_step(self, batch):
        out = self.model(self._static_batch(batch))
        loss = self.loss_fn(out)
        # torch.autograd.grad w.r.t. the per-call leaf aliases the executor made of the parameters (executor.run_plan): no
        # long-lived AccumulateGrad node (which remembers the stream it was created on) takes part in a captured backward
        alias = self._cur_spec.param_aliases
        grads = torch.autograd.grad(loss, [alias.get(id(p), p) for p in self.params], allow_unused=True)
        for p, g in zip(self.params, grads):
            p.grad = g
        return loss.detach()

    def _capture(self, batch):
        torch.cuda.synchronize(self.dev)
        loss_e, run = self._exact_step(batch)
        assert run is not None, 'GraphedStep needs a backbone running through the plan executor'
        self._make_buffers(batch, run)
        assert self._load(batch)
        # warm-up of the static path on a side stream (allocator pools, lazy kernel attributes), as torch.cuda.graphs asks for
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(max(self.warmup, 1)):
                for p in self.params:
                    p.grad = None
                self._static_step(batch)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        for p in self.params:
            p.grad = None
        self.overflow.zero_()
        self.graph = torch.cuda.CUDAGraph()
        lib = _lib.load()
        l0 = lib.vc_launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._static_step(batch)
        self.launches_per_replay = int(lib.vc_launch_count() - l0)     # this library's kernels inside one replay
        self.recaptures += 1
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------------------------------------
    def _overflowed(self):
        """Largest overflow reported by the replays that have COMPLETED so far (never blocks: the flag of a step is read
        back asynchronously and looked at when its event has fired, typically one or two calls later)."""
        worst = 0
        while self._pending and self._pending[0][0].query():
            ev, slot = self._pending.pop(0)
            worst = max(worst, int(self.ovf_host[slot]))
            if int(self.err_host[slot]) != 0:
                raise _lib.VirConvLibraryError('a tensor-core kernel\'s pipeline wait timed out inside a replayed step '
                                               '(error flag set): the results of that step are invalid')
        return worst

    def __call__(self, batch):
        """-> the step's loss (a device tensor owned by the graph: read or copy it before the next call)."""
        if self.graph is None:
            self._capture(batch)
        if self.check_overflow and self._overflowed():
            # an EARLIER batch did not fit: its gradients were computed on clamped row sets.  Grow and re-capture; the
            # caller sees the event through `recaptures` (a training loop may want to discard those optimizer steps).
            self.margin *= 1.25
            self._capture(batch)
        if not self._load(batch):
            self.margin *= 1.25
            self._capture(batch)            # more input rows than the input buffers hold
            assert self._load(batch)
        self.graph.replay()
        if self.check_overflow:
            if len(self._pending) >= self.ovf_host.numel():     # (host far ahead of the GPU: wait for the oldest read-back)
                self._pending[0][0].synchronize()
                if self._overflowed():
                    self.margin *= 1.25
                    self._capture(batch)
                    return self(batch)
            slot = self._slot
            self._slot = (slot + 1) % self.ovf_host.numel()
            self.ovf_host[slot:slot + 1].copy_(self.overflow, non_blocking=True)
            self.err_host[slot:slot + 1].copy_(ops.tc_error_flag(self.dev), non_blocking=True)
            self.overflow.zero_()
            ev = torch.cuda.Event()
            ev.record()
            self._pending.append((ev, slot))
        return self.loss


class _PipeInstance(GraphedStep):
    """One of the two buffer sets of a PipelinedStep: the step captured as TWO graphs over one caller-owned arena — the index
    graph (voxeliser, rulebooks, projections: everything that depends on coordinates only) and the feature graph (weight images,
    conv + BN + ReLU, loss, backward)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.g_idx = self.g_feat = None
        self.feat_done = None
        self.grads = None

    def _static_batch(self, batch):
        bd = super()._static_batch(batch)
        spec = bd['virconv_static']
        spec.phase, spec.arena = 0, None            # (the spec object is shared between calls when there is no voxeliser)
        return bd

    def _index_part(self, batch):
        bd = self._static_batch(batch)              # (in-graph voxeliser, if any)
        spec = bd['virconv_static']
        spec.phase, spec.arena = 1, self.arena
        self.model.index_phase(bd)
        return bd

    def _feature_part(self, bd):
        spec = bd['virconv_static']
        spec.phase = 2
        out = self.model(bd)
        loss = self.loss_fn(out)
        alias = spec.param_aliases
        grads = torch.autograd.grad(loss, [alias.get(id(p), p) for p in self.params], allow_unused=True)
        self.grads = list(grads)
        return loss.detach()

    def _capture(self, batch):
        torch.cuda.synchronize(self.dev)
        loss_e, run = self._exact_step(batch)
        assert run is not None, 'PipelinedStep needs a backbone running through the plan executor'
        self._make_buffers(batch, run)
        assert self._load(batch)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(max(self.warmup, 1)):        # unphased static steps: allocator pools, kernel attributes, arena size
                for p in self.params:
                    p.grad = None
                self._static_step(batch)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        plan = self.model._plan()
        self.arena = torch.empty(executor._arena_bytes(plan, self.cap0, self.dev, True), dtype=torch.uint8, device=self.dev)
        for p in self.params:
            p.grad = None
        self.overflow.zero_()
        lib = _lib.load()
        l0 = lib.vc_launch_count()
        self.g_idx = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_idx):
            self._bd = self._index_part(batch)
        self.g_feat = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_feat):
            self.loss = self._feature_part(self._bd)
        self.launches_per_replay = int(lib.vc_launch_count() - l0)
        self.recaptures += 1
        self.feat_done = None
        torch.cuda.synchronize(self.dev)


class PipelinedStep:
    """GraphedStep with the index work of step t+1 running BESIDE the feature work of step t.

    Consecutive replays of one monolithic graph cannot overlap, so the head of every step — voxelisation, the first rulebook,
    the strided convs' count -> indices -> tables chain — sits exposed in front of its first convolution (the exact-mode executor
    hides it behind the previous step's backward: 2.86 vs 3.01 ms per step).  Here every step is two graphs (see _PipeInstance)
    and there are two buffer sets, used alternately: the index graph of a step is launched on its own stream as soon as ITS
    buffer set is free (two steps back), the feature graph on the caller's stream once the index graph is done.  Rulebooks do not
    depend on the weights, so the overlap does not change what a training step computes.  Parameter `.grad` tensors are switched
    to the buffer set of the step that just ran."""

    def __init__(self, model, loss_fn, params=None, **kwargs):
        self.inst = [_PipeInstance(model, loss_fn, params, **kwargs) for _ in range(2)]
        self.params = self.inst[0].params
        self.dev = self.inst[0].dev
        self.s_idx = torch.cuda.Stream(device=self.dev)
        self.t = 0

    @property
    def recaptures(self):
        return sum(i.recaptures for i in self.inst)

    @property
    def launches_per_replay(self):
        return self.inst[0].launches_per_replay

    def __call__(self, batch):
        inst = self.inst[self.t & 1]
        self.t += 1
        main = torch.cuda.current_stream(self.dev)
        if inst.g_feat is None:
            inst._capture(batch)
        if inst.check_overflow and inst._overflowed():
            inst.margin *= 1.25
            inst._capture(batch)
        with torch.cuda.stream(self.s_idx):
            if inst.feat_done is not None:
                self.s_idx.wait_event(inst.feat_done)          # the buffer set's previous step (two steps back) is through
            else:
                self.s_idx.wait_stream(main)
            ok = inst._load(batch)
            if ok:
                inst.g_idx.replay()
                ev = torch.cuda.Event()
                ev.record(self.s_idx)
        if not ok:                                             # more input rows than the buffers hold: grow, re-capture, run again
            inst.margin *= 1.25
            inst._capture(batch)
            self.t -= 1
            return self(batch)
        main.wait_event(ev)
        inst.g_feat.replay()
        for p, g in zip(self.params, inst.grads):
            p.grad = g
        if inst.check_overflow:
            if len(inst._pending) >= inst.ovf_host.numel():
                inst._pending[0][0].synchronize()
            slot = inst._slot
            inst._slot = (slot + 1) % inst.ovf_host.numel()
            inst.ovf_host[slot:slot + 1].copy_(inst.overflow, non_blocking=True)
            inst.err_host[slot:slot + 1].copy_(ops.tc_error_flag(self.dev), non_blocking=True)
            inst.overflow.zero_()
            e2 = torch.cuda.Event()
            e2.record()
            inst._pending.append((e2, slot))
        inst.feat_done = torch.cuda.Event()
        inst.feat_done.record(main)
        return inst.loss
