"""View-sharded data parallelism: one process per GPU, Gaussians
replicated, camera views sharded, ONE all-reduce of a flat gradient bucket per step.

The reference is single-process/single-GPU (no collective anywhere); this is the only
cross-GPU exchange the path needs.  Backend-agnostic on purpose (NCCL on the B200 box,
gloo in the CPU tests): nothing here touches CUDA directly.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def view_for_rank(step: int, rank: int, world: int, n_views: int) -> int:
    """Round-robin: at step s rank r renders view (s*world + r) mod n_views, so a step covers
    `world` distinct consecutive views and an epoch covers all of them."""
    return (step * world + rank) % n_views


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Static partition used by render-only sweeps (C5): views rank, rank+world, ..."""
    return list(range(rank, n_views, world))


class GradBucket:
    """Flat fp32 bucket over the gradients of `params` (order fixed at construction).
    `allreduce()` sums (or averages) the bucket across ranks with a single collective and
    scatters the result back into each `.grad` in place."""

    def __init__(self, params: Sequence[torch.Tensor], average: bool = False, group=None):
        self.params = list(params)
        self.average = average
        self.group = group

    def nbytes(self) -> int:
        return sum(p.numel() for p in self.params) * 4

    def allreduce(self, async_op: bool = False):
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = _as_one_buffer(grads)          # zero-copy when the grads are views of one flat buffer
        in_place = flat is not None
        if not in_place:
            flat = _flatten_dense_tensors(grads)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

        def finish():
            if self.average:
                flat.div_(dist.get_world_size(self.group))
            if in_place:
                return
            for p, g, f in zip(self.params, grads, _unflatten_dense_tensors(flat, grads)):
                if p.grad is None:
                    p.grad = f.clone()
                else:
                    g.copy_(f)

        if async_op:
            return work, finish
        finish()
        return None


def _as_one_buffer(grads):
    """If the gradient tensors are contiguous views laid out in order in ONE storage with at most
    3 pad elements between them (the fused backward allocates them that way,
    renderer._RenderFrame.backward), return the flat 1-D view spanning them so the collective
    runs in place; else None."""
    try:
        g0 = grads[0]
        base = g0.untyped_storage().data_ptr()
        off = g0.storage_offset()
        end = off
        for g in grads:
            gap = g.storage_offset() - end
            if (g.dtype != g0.dtype or not g.is_contiguous() or g.untyped_storage().data_ptr() != base
                    or gap < 0 or gap > 3):
                return None
            end = g.storage_offset() + g.numel()
        return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(g0.untyped_storage(), off, (end - off,))
    except Exception:
        return None


def allreduce_stats(tensors: Iterable[torch.Tensor], group=None):
    """Densification statistics that are per-view in the reference (train.py:148-150:
    |grad_pos| accumulation and culling_mask counts) must be summed over ranks so that every
    replica takes the same clone/split/prune decisions."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
