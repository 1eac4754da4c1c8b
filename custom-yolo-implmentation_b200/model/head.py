"""Tail of the reference's ``Head.forward`` (src/model/head.py:86-121) on CUDA.

The reference concatenates, per level, the box and class conv outputs along channels, then flattens
and concatenates the levels along the anchor axis (every element is copied twice), and rebuilds the
anchor grid twice per call.  ``head_tail`` does the same with ONE ``yb_head_gather`` launch (each
element moved once), the cached anchor grid of ``utils.model_utils.make_anchors``, and a one-launch
backward (``yb_head_scatter``) instead of ``2 * n_levels`` strided slice copies.

The convolution towers of the head (head.py:46-62) are the dense network body and stay with the
caller: a reference ``Head`` uses this as

    box_outs = [m(f) for m, f in zip(self.box, x)]; cls_outs = [m(f) for m, f in zip(self.cls, x)]
    return head_tail(box_outs, cls_outs, self.stride)
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch

from .. import _cabi
from ..utils.model_utils import make_anchors

__all__ = ["head_tail", "gather_levels"]


def _pointer_table(tensors: Sequence[torch.Tensor]):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _check_levels(box_outs, cls_outs):
    if len(box_outs) != len(cls_outs) or not len(box_outs):
        raise ValueError("head_tail: need the same, non-zero number of box and class levels")
    if len(box_outs) > 8:
        raise ValueError("head_tail: at most 8 detection levels")
    b0 = box_outs[0]
    _cabi.require_cuda(b0, "box_outs[0]")
    n, box_ch = int(b0.shape[0]), int(b0.shape[1])
    nc = int(cls_outs[0].shape[1])
    for b, c in zip(box_outs, cls_outs):
        if b.dim() != 4 or c.dim() != 4:
            raise ValueError("head_tail: level tensors must be (N, C, H, W)")
        if b.shape[0] != n or c.shape[0] != n or b.shape[1] != box_ch or c.shape[1] != nc or b.shape[2:] != c.shape[2:]:
            raise ValueError("head_tail: inconsistent level shapes")
        if b.dtype != b0.dtype or c.dtype != b0.dtype or b.device != b0.device or c.device != b0.device:
            raise ValueError("head_tail: all levels must share dtype and device")
    return n, box_ch, nc, [int(b.shape[2]) * int(b.shape[3]) for b in box_outs]


def _copy_code(dtype) -> int:
    """The gather / scatter kernels are pure copies and only care about the element size: fp16 head outputs (the
    reference trains under fp16 autocast + GradScaler unless precision == 'bfloat16', train_model.py:240-246)
    travel as the 2-byte type."""
    return _cabi.YB_F32 if dtype == torch.float32 else _cabi.YB_BF16


class _GatherLevels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_levels, *levels):
        box_outs = [t.contiguous() for t in levels[:n_levels]]
        cls_outs = [t.contiguous() for t in levels[n_levels:]]
        n, box_ch, nc, hw = _check_levels(box_outs, cls_outs)
        dev = box_outs[0].device
        out = torch.empty(n, box_ch + nc, sum(hw), dtype=box_outs[0].dtype, device=dev)
        hw_host = (ctypes.c_int32 * n_levels)(*hw)
        if out.numel():
            with torch.cuda.device(dev):
                rc = _cabi.lib().yb_head_gather(_pointer_table(box_outs), _pointer_table(cls_outs), hw_host, n_levels,
                                                _copy_code(out.dtype), n, box_ch, nc, _cabi.ptr(out),
                                                _cabi.stream_ptr(dev))
            _cabi.check(rc, "yb_head_gather")
        ctx.meta = (n_levels, n, box_ch, nc, hw, [tuple(t.shape) for t in levels])
        return out

    @staticmethod
    def backward(ctx, grad):
        n_levels, n, box_ch, nc, hw, shapes = ctx.meta
        grad = grad.contiguous()
        grads = [torch.empty(s, dtype=grad.dtype, device=grad.device) for s in shapes]
        if grad.numel():
            hw_host = (ctypes.c_int32 * n_levels)(*hw)
            with torch.cuda.device(grad.device):
                rc = _cabi.lib().yb_head_scatter(_cabi.ptr(grad), hw_host, n_levels, _copy_code(grad.dtype), n,
                                                 box_ch, nc, _pointer_table(grads[:n_levels]),
                                                 _pointer_table(grads[n_levels:]), _cabi.stream_ptr(grad.device))
            _cabi.check(rc, "yb_head_scatter")
        return (None, *grads)


def gather_levels(box_outs: List[torch.Tensor], cls_outs: List[torch.Tensor]) -> torch.Tensor:
    """``cat([cat((b, c), 1).view(N, C, -1) for b, c in levels], 2)`` (head.py:86-87, :119) in one launch."""
    if box_outs[0].dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError(f"head_tail: dtype must be float32, bfloat16 or float16, got {box_outs[0].dtype}")
    return _GatherLevels.apply(len(box_outs), *box_outs, *cls_outs)


def head_tail(box_outs: List[torch.Tensor], cls_outs: List[torch.Tensor], stride: Sequence, offset: float = 0.5):
    """What ``Head.forward`` returns (head.py:121): ``(x (N, 4*reg_max + nc, A), anchors (2, A), strides (1, A))``."""
    x = gather_levels(box_outs, cls_outs)
    anchors, strides = make_anchors(box_outs, stride, offset)
    return x, anchors.transpose(0, 1), strides.transpose(0, 1)
