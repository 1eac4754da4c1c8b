"""Drop-in for the ``DFL`` block of the reference's ``src/model/model_blocks.py`` (:254-280) on CUDA.

The other blocks of that file (Conv, Residual, C3K, C3K2, SPPF, Attention, PSA) are the dense
network body and are outside the box-geometry hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _cabi

__all__ = ["DFL", "dfl_decode"]


def dfl_decode(box_logits: torch.Tensor, anchors=None, strides=None, reg_max: int = 16, want_ltrb: bool = True,
               box_format: str | None = None, scale_by_stride: bool = True):
    """One ``yb_dfl_decode`` call on ``(N, >=4*reg_max, A)`` head output (only the first 4*reg_max
    channels are read; a channel-sliced view of the full head output is accepted without a copy).

    Returns ``(ltrb (N,4,A) or None, box (N,4,A) or None)`` in fp32.  ``box_format``: None, "xywh", "xyxy".
    """
    _cabi.require_cuda(box_logits, "box_logits")
    if box_logits.dim() != 3 or box_logits.shape[1] < 4 * reg_max:
        raise ValueError(f"expected (N, >= {4 * reg_max}, A), got {tuple(box_logits.shape)}")
    x = box_logits.detach()
    n, _, a = x.shape
    if x.stride(2) != 1 or x.stride(1) != a:
        x = x[:, : 4 * reg_max].contiguous()
    image_stride = x.stride(0) if n > 1 else max(x.stride(0), 4 * reg_max * a)
    dev = x.device
    ltrb = torch.empty(n, 4, a, dtype=torch.float32, device=dev) if want_ltrb else None
    box = torch.empty(n, 4, a, dtype=torch.float32, device=dev) if box_format else None
    anc = st = None
    if box_format:
        anc = anchors.detach().to(device=dev, dtype=torch.float32).reshape(2, a).contiguous()
        if scale_by_stride:
            st = strides.detach().to(device=dev, dtype=torch.float32).reshape(a).contiguous()
    if n and a:
        with torch.cuda.device(dev):
            rc = _cabi.lib().yb_dfl_decode(_cabi.ptr(x), _cabi.dtype_code(x.dtype), n, reg_max, a, image_stride,
                                           _cabi.ptr(anc), _cabi.ptr(st), _cabi.ptr(ltrb), _cabi.ptr(box),
                                           1 if box_format == "xyxy" else 0, int(bool(scale_by_stride and box_format)),
                                           _cabi.stream_ptr(dev))
        _cabi.check(rc, "yb_dfl_decode")
    return ltrb, box


class DFL(nn.Module):
    """Softmax over the ``c1`` bins of each box side, then the expectation ``sum_j p_j * j``.

    Keeps the reference's frozen ``conv`` parameter (weights 0..c1-1) so that checkpoints holding
    ``dfl.conv.weight`` load unchanged; the forward pass is one fused CUDA kernel instead of a
    strided softmax plus a cuDNN 1x1 convolution."""

    def __init__(self, c1: int = 16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        x = torch.arange(c1, dtype=torch.float)
        self.conv.weight.data[:] = nn.Parameter(x.view(1, c1, 1, 1))
        self.c1 = c1

    def forward(self, x):
        b, c, a = x.shape
        if c != 4 * self.c1:
            raise RuntimeError(f"shape '[{b}, 4, {self.c1}, {a}]' is invalid for input of size {x.numel()}")
        ltrb, _ = dfl_decode(x, reg_max=self.c1)
        return ltrb.to(x.dtype)
