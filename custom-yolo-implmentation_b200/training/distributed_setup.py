"""Collectives of the hot path (reference: ``src/training/distributed_setup.py``).

The box path is per-image independent, so the batch shards across ranks with no data-path
collective; the only exchange is the small logging / normaliser reduction.  The reference issues
three scalar all-reduces plus ``.item()`` per epoch (``reduce_value`` at
src/training/train_model.py:285-288, :346-348); ``reduce_loss_stats`` does the same in ONE
all-reduce of the 8-float stats vector the loss kernel already wrote on the device.
Works with NCCL (CUDA tensors) and Gloo (CPU tensors, used by the world_size-2 tests).
"""
from __future__ import annotations

from typing import Union

import torch
import torch.distributed as dist

__all__ = ["reduce_value", "reduce_loss_stats", "shard_batch"]


def reduce_value(value: Union[float, torch.Tensor], average: bool = True):
    """Same contract as the reference's ``reduce_value`` (:28-63): sum (or mean) over all ranks, returned as a
    Python float (``value.item()``, :63) whatever was passed in; a tensor argument is reduced IN PLACE, as the
    reference does (:58-61).  Identity when not distributed (:41-42)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() < 2:
        return value
    with torch.no_grad():
        t = value if isinstance(value, torch.Tensor) else torch.tensor(float(value))
        if dist.get_backend() == "nccl" and not t.is_cuda:
            t = t.cuda()
        dist.all_reduce(t)
        if average:
            t /= dist.get_world_size()
        return t.item()


def reduce_loss_stats(stats: torch.Tensor, n_local: int, group=None, async_op: bool = False):
    """All-reduce of the loss kernel's 8-float vector ``[total, dfl_mean, cls_mean, num_matched, ...]``
    in a single message.  Returns ``[total, dfl_mean, cls_mean, num_matched_sum, n_global, 0..]``: the
    first three are global per-image means, i.e. what the reference's three
    ``reduce_value(..., average=True)`` calls give when every rank holds the same number of images
    (DistributedSampler with drop_last, src/data/data_loader.py:19-36).

    ``async_op=True`` returns ``(tensor, finish)``: the collective runs on NCCL's stream while the next
    step's kernels run on the caller's; call ``finish()`` before reading the tensor (it waits for
    the collective and turns the weighted sums into means).
    """
    v = stats.detach().clone().float()
    v[:3] *= float(n_local)
    v[4] = float(n_local)
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        work = dist.all_reduce(v, group=group, async_op=async_op)

    def finish():
        if async_op and work is not None:
            work.wait()
        v[:3] /= v[4]
        return v

    if async_op:
        return v, finish
    return finish()


def shard_batch(n_global: int, rank: int, world: int):
    """Contiguous image range ``[lo, hi)`` of ``rank`` (DistributedSampler/drop_last semantics)."""
    per = n_global // world
    return rank * per, (rank + 1) * per
