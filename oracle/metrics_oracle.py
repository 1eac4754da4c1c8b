"""CPU oracle for the validation bookkeeping (``DetectionMetrics.update``, src/training/metrics.py:68-160).
TEST INFRASTRUCTURE ONLY — nothing under ``custom-yolo-implmentation_b200/`` may import it.

Restated on whole arrays: the IoU matrix and the class-equality mask are built once, then the
prediction-ordered greedy assignment is a masked arg-max per prediction (first maximum, IoU > 0).  Pinned
against the live reference class by ``tests/golden/make_golden.py`` (``metrics_*.npz``).
"""
from __future__ import annotations

import numpy as np
import torch

from .decode_oracle import pairwise_iou_xywh


class MetricsOracle:
    def __init__(self, num_classes: int, iou_threshold: float = 0.5):
        self.nc, self.thr = num_classes, iou_threshold
        self.c = np.zeros(5, dtype=np.int64)                     # tp fp fn total_pred total_gt
        self.cls = np.zeros((4, num_classes), dtype=np.int64)    # tp fp fn gt_count

    def _bump(self, k, ids):
        for i in ids:
            if 0 <= int(i) < self.nc:
                self.cls[k, int(i)] += 1

    def update(self, pred: torch.Tensor, tgt: torch.Tensor, scores: torch.Tensor = None, score_threshold: float = 0.5):
        if pred.numel() == 0 and tgt.numel() == 0:
            return
        if scores is not None and pred.numel() > 0:
            pred = pred[scores >= score_threshold]
        p, m = (pred.shape[0] if pred.numel() else 0), (tgt.shape[0] if tgt.numel() else 0)
        if p == 0:                                                # metrics.py:89-96
            self.c[2] += m
            self._bump(2, tgt[:, 4].long().tolist()); self._bump(3, tgt[:, 4].long().tolist())
            return
        if m == 0:                                                # :98-104
            self.c[1] += p
            self._bump(1, pred[:, 4].long().tolist())
            return
        iou = pairwise_iou_xywh(pred[:, :4].float(), tgt[:, :4].float()).numpy()
        pc, tc = pred[:, 4].long().numpy(), tgt[:, 4].long().numpy()
        free = np.ones(m, dtype=bool)
        for i in range(p):
            cand = np.where(free & (tc == pc[i]) & (iou[i] > 0), iou[i], -1.0)
            j = int(cand.argmax())                                # first maximum
            if cand[j] > 0 and cand[j] >= self.thr:
                self.c[0] += 1; free[j] = False; self._bump(0, [pc[i]])
            else:
                self.c[1] += 1; self._bump(1, [pc[i]])
        self.c[2] += int(free.sum())
        self._bump(3, tc.tolist()); self._bump(2, tc[free].tolist())
        self.c[3] += p; self.c[4] += m

    def vector(self) -> np.ndarray:
        out = np.zeros(8 + 4 * self.nc, dtype=np.int64)
        out[:5] = self.c
        out[8:] = self.cls.reshape(-1)
        return out
