"""Drop-in for the hot-path parts of the reference's ``src/training/metrics.py`` on CUDA.

  box_iou_batch(boxes1, boxes2)            reference :6-41     pairwise xywh IoU, eps 1e-6
  DetectionMetrics(num_classes, iou_thr)   reference :44-207   greedy class-matched TP/FP/FN bookkeeping
  compute_average_iou(predictions, targets) reference :210-235

``DetectionMetrics.update`` is a pure-Python O(P*G) double loop with one ``.item()`` per element in the
reference, called once per image (src/training/train_model.py:326-328); here the matching runs on the
GPU (one warp per image, ``csrc/metrics.cu``), the counters live on the device and are read back only by
``compute()`` / the counter properties.  ``update_batch`` takes the device tensors of
``decode_predictions_raw`` directly, so a validation step needs no per-image host work at all.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from .. import _cabi

__all__ = ["box_iou_batch", "DetectionMetrics", "compute_average_iou"]


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
    return out.to(boxes1.dtype)


class DetectionMetrics:
    """Same constructor, ``reset`` / ``update`` / ``compute`` / ``get_class_metrics`` contract and the same
    attribute names as the reference class (``true_positives``, ``class_tp``, ...), which are read from the
    device counters on access.  The reference's early returns are kept (an image with no predictions or no
    targets does not advance ``total_predictions`` / ``total_ground_truths``, metrics.py:87-104)."""

    def __init__(self, num_classes: int, iou_threshold: float = 0.5, device=None):
        self.num_classes = num_classes
        self.iou_threshold = iou_threshold
        self._device = torch.device(device) if device is not None else None
        self._counters = None
        self.reset()

    # ---- state ----
    def reset(self):
        if self._counters is not None:
            self._counters.zero_()

    def _state(self, device) -> torch.Tensor:
        if self._counters is None:
            self._device = torch.device(device)
            self._counters = torch.zeros(8 + 4 * self.num_classes, dtype=torch.int64, device=self._device)
        return self._counters

    def _host(self) -> torch.Tensor:
        if self._counters is None:
            return torch.zeros(8 + 4 * self.num_classes, dtype=torch.int64)
        return self._counters.cpu()

    true_positives = property(lambda self: int(self._host()[0]))
    false_positives = property(lambda self: int(self._host()[1]))
    false_negatives = property(lambda self: int(self._host()[2]))
    total_predictions = property(lambda self: int(self._host()[3]))
    total_ground_truths = property(lambda self: int(self._host()[4]))

    def _class(self, k):
        nc = self.num_classes
        return self._host()[8 + k * nc: 8 + (k + 1) * nc].float()

    class_tp = property(lambda self: self._class(0))
    class_fp = property(lambda self: self._class(1))
    class_fn = property(lambda self: self._class(2))
    class_gt_count = property(lambda self: self._class(3))

    # ---- updates ----
    def update_batch(self, pred_rows: torch.Tensor, pred_count: torch.Tensor, gt, gt_offsets: torch.Tensor = None,
                     gmax: int = None, pred_scores: torch.Tensor = None, score_threshold: float = 0.5,
                     skip_empty_targets: bool = True):
        """All images of a batch in one launch.  ``pred_rows (N, K, 5)`` / ``pred_count (N,) int32`` as written by
        ``decode_predictions_raw``; ``gt (sum Mi, 5)`` / ``gt_offsets (N+1,) int32`` as built by ``pack_gt``, or a
        ``PackedGT`` (on any device: it is moved) in place of ``gt``.

        ``skip_empty_targets`` (default, = the reference's validation loop, train_model.py:326-328, which calls
        ``update`` only for images that HAVE targets): images without ground truth add nothing.  ``False`` applies
        ``update``'s own rule to them (their predictions count as false positives, metrics.py:98-104)."""
        _cabi.require_cuda(pred_rows, "pred_rows")
        n, k = pred_rows.shape[0], pred_rows.shape[1]
        dev = pred_rows.device
        if hasattr(gt, "offsets") and hasattr(gt, "counts"):            # a PackedGT
            gt_offsets, gmax, gt = gt.offsets, (max(gt.counts) if len(gt.counts) else 0), gt.gt
        if gt_offsets is None or gt_offsets.dim() != 1 or gt_offsets.numel() != n + 1:
            raise ValueError(f"gt_offsets must have shape ({n + 1},) for a batch of {n}")
        if gmax is None:
            raise ValueError("gmax (the largest number of targets of one image) is required with raw gt / gt_offsets")
        gt_offsets = gt_offsets.to(device=dev, dtype=torch.int32).contiguous()
        rows = pred_rows.detach().float().contiguous()
        cnt = pred_count.to(device=dev, dtype=torch.int32).contiguous()
        sc = None if pred_scores is None else pred_scores.detach().to(device=dev, dtype=torch.float32).contiguous()
        g = gt.detach().to(device=dev, dtype=torch.float32).contiguous()
        c = self._state(dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().yb_detection_match(_cabi.ptr(rows) if k else None, k, _cabi.ptr(cnt), _cabi.ptr(sc),
                                                float(score_threshold), _cabi.ptr(g) if g.shape[0] else None,
                                                _cabi.ptr(gt_offsets), int(gmax), n, self.num_classes,
                                                float(self.iou_threshold), int(bool(skip_empty_targets)), _cabi.ptr(c),
                                                _cabi.stream_ptr(dev))
        _cabi.check(rc, "yb_detection_match")

    def update(self, predictions: torch.Tensor, targets: torch.Tensor, pred_scores: torch.Tensor = None,
               score_threshold: float = 0.5):
        """One image: ``predictions (N, 5)`` / ``targets (M, 5)`` rows ``[x, y, w, h, class_id]`` (reference :68-160)."""
        if predictions.numel() == 0 and targets.numel() == 0:
            return
        dev = predictions.device if predictions.is_cuda else targets.device
        _cabi.require_cuda(predictions if predictions.numel() else targets, "predictions/targets")
        p = int(predictions.shape[0]) if predictions.numel() else 0
        m = int(targets.shape[0]) if targets.numel() else 0
        rows = predictions.reshape(1, p, -1)[:, :, :5] if p else torch.zeros(1, 0, 5, device=dev)
        sc = None if (pred_scores is None or p == 0) else pred_scores.reshape(1, p)
        gt = targets[:, :5] if m else torch.zeros(0, 5, device=dev)
        off = torch.tensor([0, m], dtype=torch.int32).to(dev)
        cnt = torch.tensor([p], dtype=torch.int32).to(dev)
        self.update_batch(rows, cnt, gt, off, m, sc, score_threshold, skip_empty_targets=False)

    # ---- results (same formulas as the reference, :162-207) ----
    def compute(self) -> Dict[str, float]:
        h = self._host()
        nc = self.num_classes
        tp, fp, fn = (int(h[i]) for i in range(3))
        precision = tp / (tp + fp + 1e-6)
        recall = tp / (tp + fn + 1e-6)
        f1 = 2 * (precision * recall) / (precision + recall + 1e-6)
        ctp, cfp, cgt = h[8:8 + nc].float(), h[8 + nc:8 + 2 * nc].float(), h[8 + 3 * nc:8 + 4 * nc].float()
        class_precision = ctp / (ctp + cfp + 1e-6)
        valid = cgt > 0
        m_ap = class_precision[valid].mean().item() if valid.sum() > 0 else 0.0
        return {"precision": float(precision), "recall": float(recall), "f1_score": float(f1), "mAP": float(m_ap),
                "true_positives": tp, "false_positives": fp, "false_negatives": fn,
                "total_predictions": int(h[3]), "total_ground_truths": int(h[4])}

    def get_class_metrics(self, class_id: int) -> Dict[str, float]:
        tp, fp, fn, gtc = (float(self._class(k)[class_id]) for k in range(4))
        precision = tp / (tp + fp + 1e-6)
        recall = tp / (tp + fn + 1e-6)
        f1 = 2 * (precision * recall) / (precision + recall + 1e-6)
        return {"precision": float(precision), "recall": float(recall), "f1_score": float(f1), "true_positives": int(tp),
                "false_positives": int(fp), "false_negatives": int(fn), "ground_truths": int(gtc)}


def compute_average_iou(predictions: List[torch.Tensor], targets: List[torch.Tensor]) -> float:
    """Mean over all predictions of their best IoU with the image's targets (reference :210-235)."""
    total = None
    pairs = 0
    for pred, target in zip(predictions, targets):
        if pred.numel() == 0 or target.numel() == 0:
            continue
        best = box_iou_batch(pred[:, :4], target[:, :4]).max(dim=1)[0].sum()
        total = best if total is None else total + best
        pairs += pred.size(0)
    return (float(total.item()) if total is not None else 0.0) / (pairs + 1e-6)
