"""CPU oracle for the eval half of the hot path: candidate selection + class-aware NMS.
TEST INFRASTRUCTURE ONLY — nothing under ``custom-yolo-implmentation_b200/`` may import it.

Restates ``non_max_suppression`` (``src/utils/model_utils.py:174-279``) for the flag set the
reference itself uses (``multi_label=False``, ``labels=()``; ``agnostic``/``classes`` honoured),
returning the *indices* the parity tests need, which the reference does not expose:

  select_candidates   :206, :222, :238-245   best class (first max), strict ``> conf`` filter
  order_candidates    :259                   score descending; ties -> lowest anchor index
                                             (the reference's argsort is unstable, SURVEY Q10;
                                             bit-exact tests use unique scores)
  offset boxes        :262-263               xyxy + cls * 7680 in fp32
  greedy NMS          :264                   torchvision.ops.nms — third-party; restated in plain C
                                             in ``nms_greedy.c`` (compiled by ``oracle/Makefile``)
  cap                 :265                   first ``max_det`` keeps

The reference's wall-clock abort (:212, :275-277) is NOT reproduced: it makes the
output depend on machine speed.  Parity pinning: ``tests/golden/make_golden.py`` runs the real
reference with its clock frozen and stores its output rows; ``tests/test_oracle_golden.py``
checks this file against them, and against the installed ``torchvision.ops.nms``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

MAX_WH = 7680.0      # model_utils.py:210
MAX_NMS = 30000      # model_utils.py:211
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libnms_oracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = ctypes.CDLL(so)
        _LIB.nms_greedy_sorted.restype = ctypes.c_int
        _LIB.nms_greedy_sorted.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
    return _LIB


def nms_greedy_sorted(boxes_sorted: torch.Tensor, iou_thres: float, max_det: int) -> torch.Tensor:
    """Positions (into the sorted list) of the kept boxes, at most ``max_det``."""
    b = boxes_sorted.detach().to(torch.float32).contiguous().cpu()
    n = b.shape[0]
    keep = np.empty(max(min(n, max_det), 1), dtype=np.int32)
    k = _lib().nms_greedy_sorted(b.data_ptr(), n, float(iou_thres), int(min(max_det, max(n, 0))), keep.ctypes.data)
    return torch.from_numpy(keep[:k].astype(np.int64))


def xywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    """model_utils.py:153-172 — ``x -/+ w/2`` with the half computed first."""
    dw, dh = b[..., 2] / 2, b[..., 3] / 2
    return torch.stack((b[..., 0] - dw, b[..., 1] - dh, b[..., 0] + dw, b[..., 1] + dh), -1)


@dataclass
class NmsTrace:
    rows: List[torch.Tensor]       # per image (k, 6 + nm)  [x1,y1,x2,y2,conf,cls, extras]
    keep_anchor: List[torch.Tensor]  # per image (k,) int64 anchor index of each kept row
    n_candidates: List[int]


def nms_forward(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45,
                classes: Optional[Sequence[int]] = None, agnostic: bool = False, max_det: int = 300,
                nc: int = 0, multi_label: bool = False) -> NmsTrace:
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    pred = prediction.detach().float().cpu()
    bs = pred.shape[0]
    nc = nc or (pred.shape[1] - 4)
    nm = pred.shape[1] - nc - 4
    rows, keep_anchor, ncand = [], [], []
    multi_label = multi_label and nc > 1                          # model_utils.py:213
    for i in range(bs):
        x = pred[i].transpose(0, 1)                               # (A, 4+nc+nm)
        if multi_label:
            # every (anchor, class) pair above the threshold, in nonzero() order (:240-242), then as below
            ai, cj = (x[:, 4:4 + nc] > conf_thres).nonzero(as_tuple=True)
            sc = x[ai, 4 + cj]
            if classes is not None and ai.numel():
                ok = (cj[:, None] == torch.tensor(list(classes))[None, :]).any(1)
                ai, cj, sc = ai[ok], cj[ok], sc[ok]
            ncand.append(int(ai.numel()))
            if ai.numel() == 0:
                rows.append(torch.zeros(0, 6 + nm)); keep_anchor.append(torch.zeros(0, dtype=torch.long))
                continue
            order = torch.sort(sc, descending=True, stable=True).indices[:MAX_NMS]
            ai, cj, sc = ai[order], cj[order], sc[order]
            box = xywh_to_xyxy(x[ai, :4])
            cls_f = cj.float()
            kept = nms_greedy_sorted(box + cls_f[:, None] * (0.0 if agnostic else MAX_WH), iou_thres, max_det)
            rows.append(torch.cat((box[kept], sc[kept][:, None], cls_f[kept][:, None], x[ai[kept], 4 + nc:]), 1))
            keep_anchor.append(ai[kept])
            continue
        conf, j = x[:, 4:4 + nc].max(1)                           # first max index
        cand = (conf > conf_thres).nonzero()[:, 0]                # strict, anchor order
        if classes is not None and cand.numel():
            ok = (j[cand][:, None] == torch.tensor(list(classes))[None, :]).any(1)
            cand = cand[ok]
        ncand.append(int(cand.numel()))
        if cand.numel() == 0:
            rows.append(torch.zeros(0, 6 + nm)); keep_anchor.append(torch.zeros(0, dtype=torch.long))
            continue
        order = torch.sort(conf[cand], descending=True, stable=True).indices[:MAX_NMS]
        cand = cand[order]
        box = xywh_to_xyxy(x[cand, :4])
        cls_f = j[cand].float()
        off = cls_f[:, None] * (0.0 if agnostic else MAX_WH)
        kept = nms_greedy_sorted(box + off, iou_thres, max_det)
        sel = cand[kept]
        rows.append(torch.cat((box[kept], conf[sel][:, None], cls_f[kept][:, None], x[sel, 4 + nc:]), 1))
        keep_anchor.append(sel)
    return NmsTrace(rows, keep_anchor, ncand)
