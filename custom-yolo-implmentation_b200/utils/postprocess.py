"""Fused inference post-processing: what ``Model.inference`` does after the
network (src/model/model_builder.py:123-139) — split, DFL expectation, dist2bbox, stride multiply,
concatenate, NMS — as one fused decode launch plus the batched NMS launches.
"""
from __future__ import annotations

import torch

from ..model.model_blocks import dfl_decode
from .model_utils import non_max_suppression

__all__ = ["postprocess_inference"]


def postprocess_inference(x: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor, num_classes: int,
                          conf_thres: float = 0.25, iou_thres: float = 0.45, max_det: int = 300, agnostic: bool = False,
                          classes=None, apply_sigmoid: bool = False):
    """``x (N, 64 + nc, A)`` raw head output -> list of ``(n, 6)`` ``[x1, y1, x2, y2, conf, cls]`` per image.

    ``apply_sigmoid=False`` reproduces the reference, which feeds RAW class logits to NMS as scores
    (model_builder.py:123, :136-139; SURVEY Q9 — logits above 1 then trip nothing but look odd);
    ``True`` applies the sigmoid the reference forgot.
    """
    _, box = dfl_decode(x, anchors, strides, want_ltrb=False, box_format="xywh", scale_by_stride=True)
    cls = x[:, 64:64 + num_classes].float()
    if apply_sigmoid:
        cls = cls.sigmoid()
    y = torch.cat((box, cls), 1)
    return non_max_suppression(y, conf_thres=conf_thres, iou_thres=iou_thres, classes=classes, agnostic=agnostic,
                               max_det=max_det, nc=num_classes)
