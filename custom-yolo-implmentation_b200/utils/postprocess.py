"""Fused inference post-processing: what ``Model.inference`` does after the network
(src/model/model_builder.py:123-139) — split, DFL expectation, dist2bbox, stride multiply, concatenate, NMS — through
ONE entry point of the C ABI, ``yb_postprocess``: a decode launch that leaves ``(N, 4, A)`` boxes in the workspace and the
batched NMS launches, which read their scores straight from the head output.  No torch op touches the data and the
``(N, 4 + nc, A)`` tensor the reference concatenates (:136) is never built.
"""
from __future__ import annotations

import torch

from .. import _cabi

__all__ = ["postprocess_inference", "postprocess_inference_raw"]


def postprocess_inference_raw(x: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor, num_classes: int,
                              conf_thres: float = 0.25, iou_thres: float = 0.45, max_det: int = 300, agnostic: bool = False,
                              classes=None, apply_sigmoid: bool = False, reg_max: int = 16, want_anchor: bool = False):
    """One ``yb_postprocess`` call.  Returns device tensors ``(rows (N, max_det, 6), count (N,), anchor or None)`` without
    any host synchronisation."""
    _cabi.require_cuda(x, "x")
    if x.dim() != 3 or x.shape[1] != 4 * reg_max + num_classes:
        raise ValueError(f"head output must be (N, {4 * reg_max} + {num_classes}, A), got {tuple(x.shape)}")
    head = x.detach()
    if head.dtype not in (torch.float32, torch.bfloat16):
        head = head.float()                            # fp16 autocast outputs: as the losses do
    head = head.contiguous()
    n, _, a = head.shape
    dev = head.device
    anc = anchors.detach().to(device=dev, dtype=torch.float32).contiguous()
    st = strides.detach().to(device=dev, dtype=torch.float32).contiguous()
    if anc.shape != (2, a) or st.numel() != a:
        raise ValueError(f"anchors must be (2, {a}) and strides (1, {a}); got {tuple(anc.shape)}, {tuple(st.shape)}")
    lib = _cabi.lib()
    ws = torch.empty(max(lib.yb_postprocess_workspace_bytes(n, a), 16), dtype=torch.uint8, device=dev)
    rows = torch.empty(n, max_det, 6, dtype=torch.float32, device=dev)
    count = torch.empty(n, dtype=torch.int32, device=dev)
    anchor = torch.empty(n, max_det, dtype=torch.int32, device=dev) if want_anchor else None
    filt = None
    if classes is not None:
        filt = torch.tensor([int(c) for c in classes], dtype=torch.int32).to(dev)
    with torch.cuda.device(dev):
        rc = lib.yb_postprocess(_cabi.ptr(head), _cabi.dtype_code(head.dtype), n, num_classes, reg_max, a, _cabi.ptr(anc),
                                _cabi.ptr(st), int(bool(apply_sigmoid)), float(conf_thres), float(iou_thres), int(max_det),
                                int(bool(agnostic)), _cabi.ptr(filt), 0 if filt is None else filt.numel(), _cabi.ptr(rows),
                                _cabi.ptr(count), _cabi.ptr(anchor), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "yb_postprocess")
    return rows, count, anchor


def postprocess_inference(x: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor, num_classes: int,
                          conf_thres: float = 0.25, iou_thres: float = 0.45, max_det: int = 300, agnostic: bool = False,
                          classes=None, apply_sigmoid: bool = False):
    """``x (N, 64 + nc, A)`` raw head output -> list of ``(n, 6)`` ``[x1, y1, x2, y2, conf, cls]`` per image, exactly what
    ``Model.inference`` returns (model_builder.py:139).

    ``apply_sigmoid=False`` reproduces the reference, which feeds RAW class logits to NMS as scores
    (model_builder.py:123, :136-139; SURVEY Q9 — logits above 1 then trip nothing but look odd);
    ``True`` applies the sigmoid the reference forgot, inside the scan kernel.
    """
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if classes is not None and len(classes) == 0:
        return [torch.zeros((0, 6), device=x.device)] * x.shape[0]
    rows, count, _ = postprocess_inference_raw(x, anchors, strides, num_classes, conf_thres, iou_thres, max_det, agnostic,
                                               classes, apply_sigmoid)
    counts = count.tolist()                                  # the only host sync
    empty = torch.zeros((0, 6), device=x.device)
    return [rows[i, :c] if c else empty for i, c in enumerate(counts)]
