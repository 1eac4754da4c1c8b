"""Drop-in for ``box_iou_batch`` of the reference's ``src/training/metrics.py`` (:6-41) on CUDA.

``DetectionMetrics`` (host-side TP/FP/FN bookkeeping, :44-207) is the next row of the scope table
 and is not part of this package yet.
"""
from __future__ import annotations

import torch

from .. import _cabi

__all__ = ["box_iou_batch"]


def box_iou_batch(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """Pairwise IoU of xywh boxes ``(N, 4)`` x ``(M, 4)`` -> ``(N, M)``, eps 1e-6 in the union."""
    _cabi.require_cuda(boxes1, "boxes1")
    _cabi.require_cuda(boxes2, "boxes2")
    b1 = boxes1.detach().float().contiguous()
    b2 = boxes2.detach().float().contiguous()
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty(n, m, dtype=torch.float32, device=b1.device)
    for lo in range(0, n, 65535):
        hi = min(n, lo + 65535)
        if m == 0:
            break
        with torch.cuda.device(b1.device):
            rc = _cabi.lib().yb_box_iou_batch(_cabi.ptr(b1[lo:hi]), hi - lo, _cabi.ptr(b2), m, _cabi.ptr(out[lo:hi]),
                                              _cabi.stream_ptr(b1.device))
        _cabi.check(rc, "yb_box_iou_batch")
        _cabi.count_launches(1)
    return out.to(boxes1.dtype)
