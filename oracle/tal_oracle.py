"""CPU oracle for the task-aligned (TAL) variant of the training path.  TEST INFRASTRUCTURE ONLY —
nothing under ``custom-yolo-implmentation_b200/`` may import it.

**Parity unpinned.**  The reference contains no task-aligned assigner, no CIoU and no BCE/VFL class
loss (its README only links the Ultralytics docs, README.md:115-117, which is not a
dependency, environment.yml:1-30).  BASELINE.json's north_star nevertheless names this path, so it is
specified HERE, in plain fp32 torch-on-CPU arithmetic with autograd for the gradients; the CUDA path
(``csrc/tal.cu``) is checked against this file and against the hand-computed cases in
``tests/test_tal_oracle.py``.  Results are "parity vs the in-repo oracle", never "vs the reference".

Specification (SURVEY.md §8(a'), public YOLOv8-style task-aligned learning), per image:

  decode      predicted box of anchor i: DFL expectation -> xyxy in pixels (as losses.py:155-181)
  inside      anchor centre (ax*s, ay*s) strictly inside GT j:  min(cx-x1, cy-y1, x2-cx, y2-cy) > 1e-9
  overlap     CIoU(gt_j, pred_i) clamped at 0  (eps 1e-7; alpha_v held constant in the backward)
  metric      sigmoid(cls_logit[i, c_j])**alpha * overlap**beta            (alpha 0.5, beta 6.0)
  top-k       per GT the k = 10 inside anchors with the largest metric; ties -> lowest anchor index
  conflict    an anchor chosen by several GTs goes to the one with the largest overlap; ties -> lowest GT
  targets     t_i = metric[i, j] * max_overlap_j / (max_metric_j + 1e-9)  on the anchor's class c_j,
              the maxima taken over the GT's anchors after conflict resolution
  normaliser  tss = max(sum_i t_i over the whole batch, 1)   (all-reduced SUM / world under DDP)
  losses      cls  = sum BCEWithLogits(logits, T) / tss                      over all (anchor, class)
              cls (cls_loss="vfl", the published varifocal weighting; "VFL-BCE" in the north_star)
                   = sum w * BCEWithLogits(logits, T) / tss  with  w = vfl_alpha * sigmoid(logit)**vfl_gamma
                     on background cells (differentiated) and w = T on the one positive cell of a
                     foreground anchor (label 1 even when its target score is 0)
              box  = sum_fg (1 - CIoU(pred_i, gt_j)) * t_i / tss
              dfl  = sum_fg mean_4sides[CE(left)*wl + CE(right)*wr] * t_i / tss   (target clamp [0, 14.99])
              total = lambda_box * box + lambda_cls * cls + lambda_dfl * dfl
  backward    the assignment, T and t_i are constants (as if computed under no_grad); gradients flow
              through the BCE, the CIoU of the foreground anchors and their DFL rows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from .loss_oracle import decode_boxes, dfl_loss_rows

EPS_CIOU = 1e-7
EPS_IN = 1e-9
EPS_NORM = 1e-9


def ciou(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    """Complete-IoU of xyxy boxes, broadcasting; last dim 4.  alpha_v is a constant in the backward."""
    x1, y1, x2, y2 = b1.unbind(-1)
    u1, v1, u2, v2 = b2.unbind(-1)
    w1, h1 = x2 - x1, y2 - y1 + EPS_CIOU
    w2, h2 = u2 - u1, v2 - v1 + EPS_CIOU
    inter = (torch.min(x2, u2) - torch.max(x1, u1)).clamp(0) * (torch.min(y2, v2) - torch.max(y1, v1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + EPS_CIOU
    iou = inter / union
    cw = torch.max(x2, u2) - torch.min(x1, u1)
    ch = torch.max(y2, v2) - torch.min(y1, v1)
    c2 = cw * cw + ch * ch + EPS_CIOU
    rho2 = ((u1 + u2 - x1 - x2) ** 2 + (v1 + v2 - y1 - y2) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)) ** 2
    with torch.no_grad():
        alpha = v / (v - iou + (1 + EPS_CIOU))
    return iou - (rho2 / c2 + v * alpha)


@dataclass
class TalTrace:
    total: torch.Tensor = None
    box: torch.Tensor = None
    cls: torch.Tensor = None
    dfl: torch.Tensor = None
    tss: float = 0.0                     # un-clamped sum of target scores (local)
    num_fg: int = 0
    assigned_gt: Optional[torch.Tensor] = None      # (N, A) int64, -1 = background
    target_score: Optional[torch.Tensor] = None     # (N, A) fp32 t_i (0 for background)
    topk_anchor: List[torch.Tensor] = field(default_factory=list)   # per image (M, k) int64, -1 padded
    margin: List[torch.Tensor] = field(default_factory=list)        # per image (M,) relative k-th / (k+1)-th metric gap
    grad: Optional[torch.Tensor] = None


def assign_image(xyxy: torch.Tensor, cls_logits: torch.Tensor, gt: torch.Tensor, anc_px: torch.Tensor, topk: int,
                 alpha: float, beta: float, forced_assigned: Optional[torch.Tensor] = None):
    """Task-aligned assignment of one image.  xyxy (A,4) pixels, cls_logits (A,nc), gt (M,5), anc_px (A,2).
    Returns (assigned_gt (A,) int64 with -1, target_score (A,), topk_anchor (M,k) with -1, margin (M,)).

    ``margin``: relative gap between the GT's k-th and (k+1)-th best metric — a disagreement of the CUDA path on an
    anchor of GT j is a numerical near-tie only if margin[j] is tiny.  ``forced_assigned`` (A,) skips the top-k and
    the conflict resolution and takes this anchor -> GT map instead (the later stages — target scores, losses,
    gradients — can then be checked on the CUDA path's own assignment when a near-tie went the other way)."""
    a = xyxy.shape[0]
    m = gt.shape[0]
    g = gt[:, :4].float()
    gbox = torch.stack((g[:, 0] - g[:, 2] / 2, g[:, 1] - g[:, 3] / 2, g[:, 0] + g[:, 2] / 2, g[:, 1] + g[:, 3] / 2), 1)
    gcls = gt[:, 4].long()
    d = torch.stack((anc_px[None, :, 0] - gbox[:, None, 0], anc_px[None, :, 1] - gbox[:, None, 1],
                     gbox[:, None, 2] - anc_px[None, :, 0], gbox[:, None, 3] - anc_px[None, :, 1]), 2)
    inside = d.amin(2) > EPS_IN                                                    # (M, A)
    overlap = ciou(gbox[:, None, :], xyxy[None, :, :]).clamp(min=0) * inside
    score = cls_logits.sigmoid()[:, gcls].t()                                     # (M, A)
    metric = score.pow(alpha) * overlap.pow(beta) * inside
    # top-k among the inside anchors: metric descending, ties -> lowest anchor index
    key = torch.where(inside, metric, torch.full_like(metric, -1.0))
    ranked = torch.sort(key, dim=1, descending=True, stable=True)
    order = ranked.indices[:, :topk]
    if a > topk:
        kth, nxt = ranked.values[:, topk - 1], ranked.values[:, topk]
        margin = torch.where(nxt > 0, (kth - nxt) / kth.clamp(min=1e-30), torch.ones_like(kth))
    else:
        margin = torch.ones(m)
    chosen = torch.zeros(m, a, dtype=torch.bool)
    chosen.scatter_(1, order, True)
    chosen &= inside
    topk_anchor = torch.where(chosen.gather(1, order), order, torch.full_like(order, -1))
    # conflict: largest overlap among the GTs that chose the anchor, ties -> lowest GT index
    ov = torch.where(chosen, overlap, torch.full_like(overlap, -1.0))
    best = ov.max(0)
    assigned = torch.where(chosen.any(0), best.indices, torch.full((a,), -1, dtype=torch.long))
    if forced_assigned is not None:
        assigned = forced_assigned.long()
    pos = torch.zeros(m, a, dtype=torch.bool)
    fg = assigned >= 0
    pos[assigned[fg], fg.nonzero()[:, 0]] = True
    max_metric = (metric * pos).amax(1)
    max_overlap = (overlap * pos).amax(1)
    norm = (metric * pos * (max_overlap / (max_metric + EPS_NORM))[:, None]).amax(0)    # (A,)
    return assigned, norm * fg, topk_anchor, margin


def tal_forward(preds: torch.Tensor, gts: Sequence[torch.Tensor], anchors: torch.Tensor, strides: torch.Tensor,
                num_classes: int, lambda_box: float = 1.5, lambda_cls: float = 1.0, lambda_dfl: float = 1.5,
                reg_max: int = 16, topk: int = 10, alpha: float = 0.5, beta: float = 6.0,
                tss_override: Optional[float] = None, cls_loss: str = "bce", vfl_alpha: float = 0.75,
                vfl_gamma: float = 2.0, forced_assigned: Optional[torch.Tensor] = None) -> TalTrace:
    """``tss_override`` replaces the local normaliser (DDP: the all-reduced sum / world);
    ``forced_assigned`` (N, A) replaces the assignment (see ``assign_image``)."""
    n = preds.shape[0]
    logits, _, xyxy, _ = decode_boxes(preds, anchors, strides, reg_max)
    a = xyxy.shape[1]
    cls = preds.float().transpose(1, 2)[:, :, 4 * reg_max:]
    anc = anchors.float().transpose(0, 1)
    st = strides.float().transpose(0, 1)
    anc_px = anc * st
    tr = TalTrace()
    tr.assigned_gt = torch.full((n, a), -1, dtype=torch.long)
    tr.target_score = torch.zeros(n, a)
    target = torch.zeros(n, a, num_classes)
    label = torch.zeros(n, a, num_classes)
    with torch.no_grad():
        for b in range(n):
            gt = gts[b]
            if gt.numel() == 0:
                tr.topk_anchor.append(torch.zeros(0, topk, dtype=torch.long))
                tr.margin.append(torch.zeros(0))
                continue
            asg, t, tk, mg = assign_image(xyxy[b], cls[b], gt, anc_px, topk, alpha, beta,
                                          None if forced_assigned is None else forced_assigned[b])
            tr.assigned_gt[b], tr.target_score[b] = asg, t
            tr.topk_anchor.append(tk)
            tr.margin.append(mg)
            fg = asg >= 0
            target[b, fg.nonzero()[:, 0], gt[asg[fg], 4].long()] = t[fg]
            label[b, fg.nonzero()[:, 0], gt[asg[fg], 4].long()] = 1.0
    tr.tss = float(tr.target_score.sum())
    tr.num_fg = int((tr.assigned_gt >= 0).sum())
    tss = max(tss_override if tss_override is not None else tr.tss, 1.0)
    bce = F.binary_cross_entropy_with_logits(cls, target, reduction="none")
    if cls_loss == "vfl":
        weight = vfl_alpha * cls.sigmoid().pow(vfl_gamma) * (1.0 - label) + target * label
        tr.cls = (bce * weight).sum() / tss
    else:
        tr.cls = bce.sum() / tss
    box_sum = torch.zeros(())
    dfl_sum = torch.zeros(())
    for b in range(n):
        fg = (tr.assigned_gt[b] >= 0).nonzero()[:, 0]
        if fg.numel() == 0:
            continue
        g = gts[b][tr.assigned_gt[b, fg], :4].float()
        gbox = torch.stack((g[:, 0] - g[:, 2] / 2, g[:, 1] - g[:, 3] / 2, g[:, 0] + g[:, 2] / 2, g[:, 1] + g[:, 3] / 2), 1)
        w = tr.target_score[b, fg]
        box_sum = box_sum + ((1.0 - ciou(xyxy[b, fg], gbox)) * w).sum()
        s = st[fg, 0]
        t = torch.stack((anc[fg, 0] - gbox[:, 0] / s, anc[fg, 1] - gbox[:, 1] / s, gbox[:, 2] / s - anc[fg, 0],
                         gbox[:, 3] / s - anc[fg, 1]), 1).clamp(0, reg_max - 1 - 0.01)
        z = logits[b, fg]
        rows = sum(dfl_loss_rows(z[:, k], t[:, k]) for k in range(4)) / 4.0
        dfl_sum = dfl_sum + (rows * w).sum()
    tr.box = box_sum / tss
    tr.dfl = dfl_sum / tss
    tr.total = lambda_box * tr.box + lambda_cls * tr.cls + lambda_dfl * tr.dfl
    return tr


def tal_forward_backward(preds, gts, anchors, strides, num_classes, **kw) -> TalTrace:
    leaf = preds.detach().clone().requires_grad_(True)
    tr = tal_forward(leaf, gts, anchors, strides, num_classes, **kw)
    tr.total.backward()
    tr.grad = leaf.grad.detach()
    for k in ("total", "box", "cls", "dfl"):
        setattr(tr, k, getattr(tr, k).detach())
    return tr
