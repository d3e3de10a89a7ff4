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
