"""CPU oracle for the training half of the box-geometry hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain fp32 torch-on-CPU tensor arithmetic, of what the reference
computes in ``src/model/losses.py`` — it is the checker that ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs compare and time against.  Nothing
under ``custom-yolo-implmentation_b200/`` may import it: the product path is CUDA only.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against the reference *itself*, imported from ``/root/reference`` in the build container by
``tests/golden/make_golden.py``; the resulting fixtures (``tests/golden/*.npz``) are checked by
``tests/test_oracle_golden.py`` on every run.

The oracle is staged so that each stage's integer / float outputs can be compared separately:

  decode_boxes            losses.py:155-188   DFL softmax-expectation -> ltrb -> pixel xyxy / xywh
  match_nearest_center    losses.py:211-215   cdist (ATen matmul form, K=4) + first-min argmin
  dfl_target_bins         losses.py:226-246   GT -> grid ltrb target, clamp [0, reg_max-1.01]
  dfl_loss_rows           losses.py:63-78     left/right cross-entropy, weighted
  iou_xywh_reference      losses.py:9-40      element-wise IoU *with the reference's b1_y2 slip*
  qfl_sum                 losses.py:46-57     quality focal loss over one image's (A, nc) logits
  loss_forward            losses.py:93-281    the whole thing, per image, returning intermediates
  loss_forward_backward   + autograd          d total / d preds, the fp32 gradient reference

Reference quirks reproduced on purpose: Q1 (``b1_y2 = h + cy/2``), Q2 (gradient
flows through the IoU soft target), Q3 (matmul-form distance, rounding order k=0..3 with FMA),
Q4/Q16 (duplicate matched anchor: last GT's row wins, every GT still receives d/dT), Q5 (matching
on predicted centres), Q6 (empty images count in the mean; an all-empty batch raises).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

EPS_IOU = 1e-6      # losses.py:40
EPS_LOG = 1e-12     # losses.py:53-54


# --------------------------------------------------------------------------------------------
# stage 1: decode
# --------------------------------------------------------------------------------------------
def decode_boxes(preds: torch.Tensor, anchors: torch.Tensor, strides: torch.Tensor, reg_max: int = 16):
    """DFL expectation decode of the box channels (losses.py:142-188).

    preds (N, 4R+nc, A) any float dtype; anchors (2, A) grid units; strides (1, A).
    Returns (box_logits (N,A,4,R) fp32 view, ltrb (N,A,4), xyxy (N,A,4), xywh (N,A,4)), pixels.
    """
    x = preds.float().transpose(1, 2)                        # (N, A, C)            :142
    n, a, _ = x.shape
    anc = anchors.transpose(0, 1).to(x.device)               # (A, 2)               :146
    st = strides.transpose(0, 1).to(x.device)                # (A, 1)               :147
    logits = x[:, :, : 4 * reg_max].view(n, a, 4, reg_max)   #                      :156
    prob = logits.softmax(3)                                 #                      :157
    bins = torch.arange(reg_max, dtype=x.dtype)
    ltrb = (prob * bins).sum(3)                              # (N, A, 4)            :159
    ax, ay, s = anc[None, :, 0], anc[None, :, 1], st[None, :, 0]
    x1 = (ax - ltrb[..., 0]) * s                             #                      :178
    y1 = (ay - ltrb[..., 1]) * s
    x2 = (ax + ltrb[..., 2]) * s
    y2 = (ay + ltrb[..., 3]) * s
    w, h = x2 - x1, y2 - y1                                  #                      :183-184
    cx, cy = (x1 + x2) / 2, (y1 + y2) / 2                    #                      :185-186
    xyxy = torch.stack((x1, y1, x2, y2), 2)
    xywh = torch.stack((cx, cy, w, h), 2)                    #                      :188
    return logits, ltrb, xyxy, xywh


# --------------------------------------------------------------------------------------------
# stage 2: nearest-predicted-centre matching
# --------------------------------------------------------------------------------------------
def _fma32(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """fp32 fused multiply-add, emulated through fp64 (a*b is exact in fp64)."""
    return (a.double() * b.double() + c.double()).float()


def center_distance(gt_xy: torch.Tensor, pred_xy: torch.Tensor, impl: str = "spec") -> torch.Tensor:
    """(M, A) Euclidean distances as ``torch.cdist`` produces them for A > 25 (losses.py:214).

    ATen evaluates ``|g|^2 + |p|^2 - 2 g.p`` as ONE K=4 fp32 GEMM row
    ``[-2gx, -2gy, |g|^2, 1] . [px, py, 1, |p|^2]`` (``_euclidean_dist``; the source is torch's, not
    under /root/reference — torch pinned 2.9.1 by environment.yml:28), then ``clamp_min(0).sqrt()``.

    impl="spec":  the stated arithmetic — norms as fl(fl(x^2)+fl(y^2)), dot accumulated k=0..3 with
                  fused multiply-adds, correctly rounded sqrt.  Probed in the build container:
                  bit-identical to ``torch.cdist`` *before* the sqrt on MKL/AVX-512; the vectorised
                  CPU ``sqrt`` of this torch build is faithfully- but not correctly-rounded (0.56 %
                  of inputs are 1 ulp off), which only matters for exact ties.
    impl="torch": literally ``torch.cdist`` (what the reference executes on this machine).
    """
    if impl == "torch":
        return torch.cdist(gt_xy, pred_xy)
    gx, gy = gt_xy[:, 0:1], gt_xy[:, 1:2]                    # (M, 1)
    px, py = pred_xy[None, :, 0], pred_xy[None, :, 1]        # (1, A)
    gn = gx * gx + gy * gy
    pn = px * px + py * py
    acc = (-2.0 * gx) * px
    acc = _fma32(-2.0 * gy, py, acc)
    acc = acc + gn                                           # fma(gn, 1, acc)
    acc = acc + pn                                           # fma(1, pn, acc)
    return acc.clamp_min(0).double().sqrt().float()          # fp64 sqrt -> fp32 is correctly rounded


def match_nearest_center(gt_xy: torch.Tensor, pred_xy: torch.Tensor, impl: str = "spec"):
    """One matched anchor per GT: first index of the minimum distance (losses.py:215).

    Returns (idx (M,) int64, best (M,) fp32, margin (M,) fp32 = second-best minus best distance).
    """
    d = center_distance(gt_xy, pred_xy, impl)
    idx = d.argmin(dim=1)
    if d.shape[1] > 1:
        two = d.topk(2, dim=1, largest=False).values
        margin = two[:, 1] - two[:, 0]
    else:
        margin = torch.full((d.shape[0],), float("inf"))
    return idx, d.gather(1, idx[:, None])[:, 0], margin


# --------------------------------------------------------------------------------------------
# stage 3: per-GT terms
# --------------------------------------------------------------------------------------------
def dfl_target_bins(gt_xywh: torch.Tensor, anc: torch.Tensor, st: torch.Tensor, reg_max: int = 16):
    """Continuous ltrb target in grid units of the matched anchor (losses.py:226-246)."""
    gx1 = gt_xywh[:, 0] - gt_xywh[:, 2] / 2
    gy1 = gt_xywh[:, 1] - gt_xywh[:, 3] / 2
    gx2 = gt_xywh[:, 0] + gt_xywh[:, 2] / 2
    gy2 = gt_xywh[:, 1] + gt_xywh[:, 3] / 2
    s = st[:, 0]
    t = torch.stack((anc[:, 0] - gx1 / s, anc[:, 1] - gy1 / s, gx2 / s - anc[:, 0], gy2 / s - anc[:, 1]), 1)
    return t.clamp(0, reg_max - 1 - 0.01)


def dfl_loss_rows(logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """(M,) left/right-bin cross-entropy for one side (losses.py:63-78, before the mean)."""
    left = target.long()
    right = left + 1
    w_left = right.float() - target
    w_right = target - left.float()
    logp = F.log_softmax(logits, dim=1)
    return -(logp.gather(1, left[:, None])[:, 0] * w_left + logp.gather(1, right[:, None])[:, 0] * w_right)


def iou_xywh_reference(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """Element-wise IoU of xywh pairs as the reference computes it (losses.py:9-40).

    box1's bottom edge is ``h + cy/2`` (losses.py:20) — kept on purpose; box2 is normal.
    """
    ax1 = box1[:, 0] - box1[:, 2] / 2
    ay1 = box1[:, 1] - box1[:, 3] / 2
    ax2 = box1[:, 0] + box1[:, 2] / 2
    ay2 = box1[:, 3] + box1[:, 1] / 2          # sic
    bx1 = box2[:, 0] - box2[:, 2] / 2
    by1 = box2[:, 1] - box2[:, 3] / 2
    bx2 = box2[:, 0] + box2[:, 2] / 2
    by2 = box2[:, 1] + box2[:, 3] / 2
    iw = (torch.min(ax2, bx2) - torch.max(ax1, bx1)).clamp(min=0)
    ih = (torch.min(ay2, by2) - torch.max(ay1, by1)).clamp(min=0)
    inter = iw * ih
    union = (ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter
    return inter / (union + EPS_IOU)


def qfl_sum(cls_logits: torch.Tensor, target: torch.Tensor, beta: float = 2.0) -> torch.Tensor:
    """Quality focal loss of one image, summed over (A, nc) and divided by A (losses.py:51-56)."""
    p = cls_logits.sigmoid()
    pos = target * (1 - p).pow(beta) * torch.log(p + EPS_LOG)
    neg = (1 - target) * p.pow(beta) * torch.log(1 - p + EPS_LOG)
    return -(pos + neg).sum() / cls_logits.shape[0]


# --------------------------------------------------------------------------------------------
# the whole loss
# --------------------------------------------------------------------------------------------
@dataclass
class LossTrace:
    """Everything the parity tests look at."""
    total: torch.Tensor
    dfl_mean: torch.Tensor
    cls_mean: torch.Tensor
    idx: List[torch.Tensor] = field(default_factory=list)        # per image (Mi,) int64 matched anchor
    margin: List[torch.Tensor] = field(default_factory=list)     # per image (Mi,) runner-up gap
    iou: List[torch.Tensor] = field(default_factory=list)        # per image (Mi,) soft target
    bins_left: List[torch.Tensor] = field(default_factory=list)  # per image (Mi,4) int64 left DFL bin
    fg_mask: Optional[torch.Tensor] = None                       # (N, A) bool: anchor matched by >=1 GT
    dfl_per_image: Optional[torch.Tensor] = None                 # (N,)
    cls_per_image: Optional[torch.Tensor] = None                 # (N,)
    grad: Optional[torch.Tensor] = None                          # (N, C, A) d total / d preds


def loss_forward(preds: torch.Tensor, gts: Sequence[torch.Tensor], anchors: torch.Tensor, strides: torch.Tensor,
                 num_classes: int, lambda_cls: float = 1.0, lambda_dfl: float = 1.5, reg_max: int = 16,
                 match_impl: str = "spec", forced_idx: Optional[Sequence[torch.Tensor]] = None) -> LossTrace:
    """Restatement of ``YoloDFLQFLoss.forward`` (losses.py:93-281).

    ``forced_idx`` replaces stage 2 (used to compare the remaining stages when a near-tie in the
    distance matrix resolves differently on two machines).  ``lambda_box`` does not appear: the
    reference stores it and never reads it (losses.py:88, :275).
    """
    n = preds.shape[0]
    logits, _, _, xywh = decode_boxes(preds, anchors, strides, reg_max)
    a = xywh.shape[1]
    cls = preds.float().transpose(1, 2)[:, :, 4 * reg_max:]      # (N, A, nc)
    anc = anchors.transpose(0, 1).to(cls.dtype)
    st = strides.transpose(0, 1).to(cls.dtype)

    tr = LossTrace(total=None, dfl_mean=None, cls_mean=None)
    tr.fg_mask = torch.zeros(n, a, dtype=torch.bool)
    dfl_img, cls_img = [], []
    any_gt = False
    for b in range(n):
        gt = gts[b]
        target = torch.zeros_like(cls[b])                        #                      :204
        if gt.numel() > 0:
            any_gt = True
            g = gt[:, :4].float()
            if forced_idx is not None:
                idx = forced_idx[b].long()
                margin = torch.full((g.shape[0],), float("nan"))
            else:
                idx, _, margin = match_nearest_center(g[:, :2], xywh[b, :, :2].detach(), match_impl)
            t = dfl_target_bins(g, anc[idx], st[idx], reg_max)
            z = logits[b, idx]                                   # (M, 4, R)
            rows = sum(dfl_loss_rows(z[:, k], t[:, k]).mean() for k in range(4)) / 4.0   # :249-252
            iou = iou_xywh_reference(xywh[b, idx], g)            #                      :256
            rows_t = torch.zeros(g.shape[0], num_classes)
            rows_t.scatter_(1, gt[:, 4].long()[:, None], iou[:, None])
            # `target_scores[idx] = matched_targets` (:261).  With duplicate anchors index_put_ is
            # order-undefined in torch (multi-threaded CPU and CUDA runs really do differ); the oracle
            # fixes the sequential meaning: the LAST GT's whole row wins (what one thread produces).
            later_same = (idx[None, :] == idx[:, None]).triu(1).any(1)      # a later GT shares my anchor
            win = ~later_same
            target[idx[win]] = rows_t[win]                                  # unique indices: deterministic
            dfl_img.append(rows)
            tr.idx.append(idx); tr.margin.append(margin); tr.iou.append(iou.detach())
            tr.bins_left.append(t.long())
            tr.fg_mask[b, idx] = True
        else:
            dfl_img.append(torch.zeros(()))
            tr.idx.append(torch.zeros(0, dtype=torch.long)); tr.margin.append(torch.zeros(0))
            tr.iou.append(torch.zeros(0)); tr.bins_left.append(torch.zeros(0, 4, dtype=torch.long))
        q = qfl_sum(cls[b], target)                              #                      :264
        if gt.numel() > 0 and bool(later_same.any()):
            # Q16: index_put_'s backward hands EVERY source row the gradient of its destination row,
            # overwritten or not.  A zero-valued term reproduces that gradient for the losing GTs.
            lose = later_same
            p = cls[b][idx[lose]].detach().sigmoid()
            dq_dt = -((1 - p).pow(2) * torch.log(p + EPS_LOG) - p.pow(2) * torch.log(1 - p + EPS_LOG)) / cls[b].shape[0]
            q = q + (dq_dt * (rows_t[lose] - rows_t[lose].detach())).sum()
        cls_img.append(q)
    if not any_gt:
        # the reference dies here with AttributeError on a python float
        raise AttributeError("'float' object has no attribute 'detach'")
    tr.dfl_per_image = torch.stack(dfl_img)
    tr.cls_per_image = torch.stack(cls_img)
    tr.dfl_mean = tr.dfl_per_image.sum() / n                     # N counts empty images    :271
    tr.cls_mean = tr.cls_per_image.sum() / n
    tr.total = lambda_dfl * tr.dfl_mean + lambda_cls * tr.cls_mean
    return tr


def loss_forward_backward(preds: torch.Tensor, gts, anchors, strides, num_classes: int, **kw) -> LossTrace:
    """Forward plus ``d total / d preds`` by autograd (fp32 reference gradient, input dtype out)."""
    leaf = preds.detach().clone().requires_grad_(True)
    tr = loss_forward(leaf, gts, anchors, strides, num_classes, **kw)
    tr.total.backward()
    tr.grad = leaf.grad.detach()
    tr.total = tr.total.detach(); tr.dfl_mean = tr.dfl_mean.detach(); tr.cls_mean = tr.cls_mean.detach()
    tr.dfl_per_image = tr.dfl_per_image.detach(); tr.cls_per_image = tr.cls_per_image.detach()
    return tr
