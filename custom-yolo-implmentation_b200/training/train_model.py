"""Drop-in for ``decode_predictions`` of the reference's ``src/training/train_model.py`` (:14-142).

The epoch loop of that file (``train``) is orchestration and stays with the reference; it calls this
function at :323 and the loss at :245 / :315.
"""
from __future__ import annotations

import torch

from .. import _cabi

__all__ = ["decode_predictions", "decode_predictions_raw"]


def decode_predictions_raw(preds, anchors, strides, conf_threshold=0.25, top_k=100, num_classes=171,
                           want_anchor: bool = False):
    """One ``yb_val_decode`` call; returns device tensors ``(rows (N, top_k, 5), count (N,), anchor)``."""
    _cabi.require_cuda(preds, "preds")
    reg_max = 16                                     # hard-coded in the reference (:33)
    n, c, a = preds.shape
    if c != 4 * reg_max + num_classes:
        raise ValueError(f"preds has {c} channels, expected 64 + {num_classes}")
    x = preds.detach().contiguous()
    dev = x.device
    anc = anchors.detach().to(device=dev, dtype=torch.float32).reshape(2, a).contiguous()
    st = strides.detach().to(device=dev, dtype=torch.float32).reshape(a).contiguous()
    lib = _cabi.lib()
    ws = torch.empty(max(lib.yb_val_decode_workspace_bytes(n, a), 16), dtype=torch.uint8, device=dev)
    rows = torch.empty(n, top_k, 5, dtype=torch.float32, device=dev)
    count = torch.empty(n, dtype=torch.int32, device=dev)
    anchor = torch.empty(n, top_k, dtype=torch.int32, device=dev) if want_anchor else None
    with torch.cuda.device(dev):
        rc = lib.yb_val_decode(_cabi.ptr(x), _cabi.dtype_code(x.dtype), n, num_classes, reg_max, a, _cabi.ptr(anc),
                               _cabi.ptr(st), float(conf_threshold), int(top_k), _cabi.ptr(rows), _cabi.ptr(count),
                               _cabi.ptr(anchor), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "yb_val_decode")
    return rows, count, anchor


def decode_predictions(preds, anchors, strides, conf_threshold=0.25, top_k=100, num_classes=171):
    """Raw head output -> per-image ``(Mi, 5)`` ``[x, y, w, h, class_id]`` (reference :14-142).

    DFL expectation decode to pixel xywh, sigmoid, best class, ``>= conf_threshold``, at most
    ``top_k`` rows per image: in anchor order when no more than ``top_k`` pass, else by descending
    score.  Two launches for the whole batch and one device-to-host copy of the counts.
    """
    rows, count, _ = decode_predictions_raw(preds, anchors, strides, conf_threshold, top_k, num_classes)
    counts = count.tolist()
    empty = torch.zeros(0, 5, device=preds.device)
    return [rows[i, :c] if c else empty for i, c in enumerate(counts)]
