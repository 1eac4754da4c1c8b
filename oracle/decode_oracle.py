"""CPU oracle for the decode / IoU-utility rows of the hot path.  TEST INFRASTRUCTURE ONLY —
nothing under ``custom-yolo-implmentation_b200/`` may import it.

  make_anchor_grid    src/utils/model_utils.py:18-70      cell-centre grid + per-anchor stride
  dfl_expectation     src/model/model_blocks.py:278-280   softmax over the 16 bins, dot with 0..15
  ltrb_to_box         src/utils/model_utils.py:120-129    dist2bbox
  val_decode          src/training/train_model.py:14-142  decode_predictions (decode + conf filter + top-k)
  pairwise_iou_xyxy   src/utils/model_utils.py:131-151    box_iou, eps 1e-7
  pairwise_iou_xywh   src/training/metrics.py:6-41        box_iou_batch, eps 1e-6

Pinned against the live reference by ``tests/golden/make_golden.py`` /
``tests/test_oracle_golden.py`` (the reference has no tests of its own).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import torch


def make_anchor_grid(shapes: Sequence[Tuple[int, int]], strides: Sequence[float], offset: float = 0.5,
                     dtype: torch.dtype = torch.float32):
    """Returns (anchors (A, 2), strides (A, 1)); x fastest within a level, levels concatenated."""
    pts, sts = [], []
    for (h, w), s in zip(shapes, strides):
        xs = torch.arange(w, dtype=dtype) + offset
        ys = torch.arange(h, dtype=dtype) + offset
        pts.append(torch.stack((xs.repeat(h), ys.repeat_interleave(w)), 1))
        sts.append(torch.full((h * w, 1), s, dtype=dtype))
    return torch.cat(pts), torch.cat(sts)


def head_tail(box_outs: Sequence[torch.Tensor], cls_outs: Sequence[torch.Tensor]) -> torch.Tensor:
    """Tail of Head.forward: per level ``cat((box, cls), 1)`` (src/model/head.py:86-87), then
    ``cat([i.view(N, no, -1) for i in x], 2)`` (:119).  Returns (N, box_ch + nc, sum H*W)."""
    n = box_outs[0].shape[0]
    levels = [torch.cat((b, c), dim=1) for b, c in zip(box_outs, cls_outs)]
    return torch.cat([lv.reshape(n, lv.shape[1], -1) for lv in levels], dim=2)


def dfl_expectation(box_logits: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """(b, 4R, a) -> (b, 4, a) expected bin per side, computed in the input dtype."""
    b, _, a = box_logits.shape
    prob = box_logits.view(b, 4, reg_max, a).softmax(2)
    bins = torch.arange(reg_max, dtype=torch.float32).to(box_logits.dtype).view(1, 1, reg_max, 1)
    return (prob * bins).sum(2)


def ltrb_to_box(dist: torch.Tensor, anchor_points: torch.Tensor, xywh: bool = True, dim: int = -1):
    lt, rb = dist.split(2, dim)
    lo, hi = anchor_points - lt, anchor_points + rb
    if xywh:
        return torch.cat(((lo + hi) / 2, hi - lo), dim)
    return torch.cat((lo, hi), dim)


@dataclass
class DecodeTrace:
    rows: List[torch.Tensor]      # per image (k, 5) [cx, cy, w, h, cls]
    anchor: List[torch.Tensor]    # per image (k,) int64 anchor index of each row
    score: List[torch.Tensor]     # per image (k,) the sigmoid score that ranked it


def val_decode(preds: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor, conf_threshold: float = 0.25,
               top_k: int = 100, num_classes: int = 171, reg_max: int = 16) -> DecodeTrace:
    """decode_predictions: xywh*stride decode, sigmoid, best class, ``>= conf``, top-k by score.

    Row order: anchor order when at most ``top_k`` candidates survive, otherwise ``torch.topk``
    order (score descending; ties -> lowest anchor index here, unspecified in the reference).
    """
    n = preds.shape[0]
    ltrb = dfl_expectation(preds[:, : 4 * reg_max, :], reg_max).permute(0, 2, 1)      # (N, A, 4)
    anc = anchors.transpose(0, 1).unsqueeze(0)
    st = strides.transpose(0, 1).unsqueeze(0)
    box = ltrb_to_box(ltrb, anc, xywh=True, dim=2) * st
    out = DecodeTrace([], [], [])
    for b in range(n):
        sc = preds[b, 4 * reg_max:, :].transpose(0, 1).sigmoid()
        best, cid = sc.max(1)
        cand = (best >= conf_threshold).nonzero()[:, 0]
        if cand.numel() > top_k:
            order = torch.sort(best[cand], descending=True, stable=True).indices[:top_k]
            cand = cand[order]
        out.anchor.append(cand)
        out.score.append(best[cand])
        if cand.numel() == 0:
            out.rows.append(torch.zeros(0, 5))
        else:
            out.rows.append(torch.cat((box[b, cand], cid[cand].float()[:, None]), 1))
    return out


def pairwise_iou_xyxy(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    lo = torch.max(b1[:, None, :2], b2[None, :, :2])
    hi = torch.min(b1[:, None, 2:], b2[None, :, 2:])
    inter = (hi - lo).clamp(0).prod(2)
    a1 = (b1[:, 2:] - b1[:, :2]).prod(1)
    a2 = (b2[:, 2:] - b2[:, :2]).prod(1)
    return inter / (a1[:, None] + a2[None, :] - inter + eps)


def _corners(b: torch.Tensor) -> torch.Tensor:
    return torch.stack((b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), 1)


def pairwise_iou_xywh(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    c1, c2 = _corners(b1), _corners(b2)
    lo = torch.max(c1[:, None, :2], c2[None, :, :2])
    hi = torch.min(c1[:, None, 2:], c2[None, :, 2:])
    wh = (hi - lo).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    a1 = (c1[:, 2] - c1[:, 0]) * (c1[:, 3] - c1[:, 1])
    a2 = (c2[:, 2] - c2[:, 0]) * (c2[:, 3] - c2[:, 1])
    return inter / (a1[:, None] + a2[None, :] - inter + 1e-6)
