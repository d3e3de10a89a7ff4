"""Data parallelism for the VirConv backbone: scenes are independent units, so the path shards with NO
data-path collective; the only exchange is the gradient all-reduce after backward (`tools/train.py:140-141` wraps the
model in DistributedDataParallel for exactly this).  One process per GPU, `torch.distributed` (NCCL over NVLink on
the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_scene_ids(step: int, rank: int, world_size: int, scenes_per_gpu: int, base: int = 0):
    """Scene ids rank `rank` processes at `step`: a disjoint, gap-free partition of
    [base + step*W*S, base + (step+1)*W*S) across ranks (the reference's DistributedSampler does the same job)."""
    start = base + (step * world_size + rank) * scenes_per_gpu
    return list(range(start, start + scenes_per_gpu))


def _as_one_buffer(grads):
    """If the gradient tensors are back-to-back fp32 slices of one storage, return that span as a 1-D tensor, else None."""
    g0 = grads[0]
    if any(g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
        return None
    st = g0.untyped_storage()
    ptr = g0.data_ptr()
    for g in grads:
        if g.untyped_storage().data_ptr() != st.data_ptr() or g.data_ptr() != ptr:
            return None
        ptr += g.numel() * 4
    flat = torch.empty(0, dtype=torch.float32, device=g0.device)
    flat.set_(st, g0.storage_offset(), ((ptr - g0.data_ptr()) // 4,), (1,))
    return flat


class PeerAllReduce:
    """One-kernel all-reduce of a flat fp32 gradient buffer over NVLink peer memory (csrc/allreduce.cu): every rank publishes
    its gradients in a symmetric buffer (torch.distributed._symmetric_memory: mapped into all peers of the node), signals,
    and sums all peers' copies straight out of their memory.  Bit-identical to NCCL's result at 2 ranks; measured back to back
    on the backbone's 1.7 MB message it is NOT faster than NCCL 2.28 on the NVSwitch box (24.5 vs 18.2 us at 2 GPUs,
    profiles/check_peer_allreduce_n2.txt) — what bench.py reports as all-reduce time (84 us at 2 GPUs, 176 us at 8) is mostly
    the skew between the ranks' backward passes, which no collective removes — so it is opt-in (VIRCONV_PEER_ALLREDUCE=1).
    Construction is collective (every rank, same `numel`)."""

    def __init__(self, numel: int, device, group=None):
        import ctypes
        import numpy as np
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel = int(numel)
        self.n_pad = (self.numel + 3) // 4 * 4
        self.buf = symm.empty(2 * self.n_pad, dtype=torch.float32, device=device)
        self.flags = symm.empty(int(self.lib.vc_allreduce_peer_flag_words(self.world)), dtype=torch.int32, device=device)
        self.flags.zero_()
        hb, hf = symm.rendezvous(self.buf, group), symm.rendezvous(self.flags, group)
        self._handles = (hb, hf)          # (keep the mappings alive)
        self.buf_ptrs = np.array([int(p) for p in hb.buffer_ptrs], dtype=np.uint64)
        self.flag_ptrs = np.array([int(p) for p in hf.buffer_ptrs], dtype=np.uint64)
        self.epoch = 0
        torch.cuda.synchronize(device)
        dist.barrier(group)               # every rank's flags are zero before anybody signals

    def __call__(self, flat: torch.Tensor, average: bool = True):
        from . import _lib, ops
        assert flat.dtype == torch.float32 and flat.is_contiguous() and flat.numel() == self.numel
        self.epoch += 1
        _lib.check(self.lib.vc_allreduce_peer_f32(self.buf_ptrs.ctypes.data, self.flag_ptrs.ctypes.data, self.rank, self.world,
                                                  flat.data_ptr(), self.numel, self.n_pad, self.epoch & 0xffffffff,
                                                  (1.0 / self.world) if average else 1.0, ops.tc_error_flag(flat.device).data_ptr(),
                                                  ops._stream()), 'vc_allreduce_peer_f32')
        return flat.numel() * 4


_PEER = {}      # (device index, numel, group id) -> PeerAllReduce, or None where symmetric memory is not available


def _peer_allreduce(flat, group):
    import os
    if os.environ.get('VIRCONV_PEER_ALLREDUCE', '0') != '1' or dist.get_backend(group) != 'nccl':
        return None
    key = (flat.device.index, flat.numel(), id(group))
    if key not in _PEER:
        peer = None
        try:
            peer = PeerAllReduce(flat.numel(), flat.device, group)
        except Exception as ex:          # noqa: BLE001   (no P2P / symmetric memory on this box: NCCL does the job)
            import sys
            sys.stderr.write(f'virconv_b200.parallel: peer-memory all-reduce unavailable ({type(ex).__name__}: {ex}); using NCCL\n')
        # all ranks or none (a rank on its own in the peer kernel would wait for signals that never come)
        ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=flat.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        _PEER[key] = peer if int(ok.item()) == 1 else None
    return _PEER[key]


def allreduce_gradients(params, average: bool = True, group=None) -> int:
    """ONE all-reduce for all gradients: flatten (one kernel), reduce, scatter back into the existing .grad tensors.
    The whole VirConv-L backbone is 1.7 MB of fp32 gradients: a latency-bound message, so a single bucket is optimal
    (SURVEY §5 comm row).  Returns the number of bytes reduced.  No-op for world_size 1 / uninitialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    flat = _as_one_buffer(grads)
    if flat is not None:
        # the plan executor writes every gradient into ONE flat buffer (executor.PlanFn.backward): reduce it in place,
        # no flatten / scatter-back kernels
        # (NCCL averages inside the collective: no separate divide kernel; gloo — the CPU tests — has no AVG)
        peer = _peer_allreduce(flat, group)
        if peer is not None:
            # one kernel over NVLink peer memory (every rank took the same branch: construction above is collective)
            return peer(flat, average)
        if average and dist.get_backend(group) == 'nccl':
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(world)
        return flat.numel() * flat.element_size()
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
    return flat.numel() * flat.element_size()
