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

__all__ = ["reduce_value", "reduce_loss_stats", "shard_batch", "PeerExchange", "enable_peer_exchange", "default_peer_exchange"]


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


# ----------------------------------------------------------------------------------------------
# peer exchange of the task-aligned normaliser (csrc/peer.cu): NVLink stores + a poll instead of a collective call
# ----------------------------------------------------------------------------------------------
class PeerExchange:
    """Peer-mapped mailboxes for the one exchange step of the task-aligned loss (the all-reduce of
    ``[sum of target scores, #foreground]`` between assignment and loss).

    Set-up is collective: every rank of ``group`` must construct it.  Each rank allocates a 2 KB mailbox with plain
    ``cudaMalloc`` (the C ABI does it: CUDA IPC cannot export memory of PyTorch's caching allocator), the 64-byte IPC
    handles travel through ``torch.distributed.all_gather_object`` and every rank maps all the others.  After that the
    exchange needs no host call at all: ``yb_tal_assign``'s last kernel stores into all mailboxes, ``yb_tal_loss``'s first
    kernel polls the own one.  Single node only (the ranks must be able to map each other's memory); the NCCL
    all-reduce of ``fused_tal_loss`` remains the general path."""

    def __init__(self, group=None, device=None):
        import ctypes

        from .. import _cabi
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _cabi.PEER_MAX_WORLD:
            raise ValueError(f"PeerExchange supports at most {_cabi.PEER_MAX_WORLD} ranks (one NVSwitch domain), got {self.world}")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lib = _cabi.lib()
        self._lib, self._own, self._opened = lib, ctypes.c_void_p(), []
        with torch.cuda.device(self.device):
            _cabi.check(lib.yb_peer_mailbox_alloc(ctypes.byref(self._own)), "yb_peer_mailbox_alloc")
            handle = ctypes.create_string_buffer(_cabi.PEER_HANDLE_BYTES)
            _cabi.check(lib.yb_peer_mailbox_export(self._own, handle), "yb_peer_mailbox_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._struct = _cabi.PeerExchangeStruct()
            self._struct.world, self._struct.rank, self._struct.seq = self.world, self.rank, 0
            for r, h in enumerate(handles):
                if r == self.rank:
                    self._struct.mailbox[r] = self._own.value
                    continue
                peer = ctypes.c_void_p()
                _cabi.check(lib.yb_peer_mailbox_open(ctypes.create_string_buffer(h, _cabi.PEER_HANDLE_BYTES), ctypes.byref(peer)),
                            f"yb_peer_mailbox_open (rank {r})")
                self._opened.append(peer)
                self._struct.mailbox[r] = peer.value
        dist.barrier(group=group)                          # nobody stores before everybody has mapped

    def next_step(self):
        """The struct for the next exchange (all ranks call this once per step, in lockstep)."""
        self._struct.seq = (self._struct.seq % 0xFFFFFFFF) + 1
        return self._struct

    def close(self):
        for p in self._opened:
            self._lib.yb_peer_mailbox_close(p)
        self._opened = []
        if self._own:
            self._lib.yb_peer_mailbox_free(self._own)
            self._own = None


_default_exchange = None


def enable_peer_exchange(group=None):
    """Make ``YoloDFLQFLoss(assigner="tal")`` / ``fused_tal_loss`` exchange the normaliser through peer mailboxes
    instead of an NCCL all-reduce.  Collective: call on every rank, after ``init_process_group``."""
    global _default_exchange
    _default_exchange = PeerExchange(group)
    return _default_exchange


def default_peer_exchange():
    return _default_exchange
