"""Drop-in for the hot-path functions of the reference's ``src/utils/model_utils.py`` on CUDA.

  make_anchors(x, strides, offset=0.5)                      reference :18-70
  dist2bbox(distance, anchor_points, xywh=True, dim=-1)     reference :120-129
  box_iou(box1, box2, eps=1e-7)                             reference :131-151
  xywh2xyxy(x)                                              reference :153-172
  non_max_suppression(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels,
                      max_det, nc)                          reference :174-279

``autopad`` / ``fuse_conv`` (conv plumbing, reference :9-16, :72-118) are outside the box-geometry
path and are not provided.  Everything here launches the CUDA kernels in ``csrc/`` through the C
ABI; inputs must live on a CUDA device.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from .. import _cabi

__all__ = ["make_anchors", "dist2bbox", "box_iou", "xywh2xyxy", "non_max_suppression", "batched_nms_raw"]

_anchor_cache = {}


def make_anchors(x: List[torch.Tensor], strides: Sequence, offset: float = 0.5):
    """Anchor cell centres ``(A, 2)`` and per-anchor stride ``(A, 1)`` in ``x[0]``'s dtype/device.

    The grid depends only on the feature-map shapes, so it is built once per
    ``(shapes, strides, offset, dtype, device)`` and cached (the reference rebuilds it twice per
    forward).  Values are produced exactly as the reference does: ``arange`` in the
    target dtype plus ``offset`` in that dtype (bf16 grids wider than 256 cells round, Q13).
    """
    assert x is not None
    dtype, device = x[0].dtype, x[0].device
    _cabi.require_cuda(x[0], "x[0]")
    shapes = tuple((int(t.shape[-2]), int(t.shape[-1])) for t in x[: len(strides)])
    key = (shapes, tuple(float(s) for s in strides), float(offset), dtype, device)
    hit = _anchor_cache.get(key)
    if hit is None:
        a = sum(h * w for h, w in shapes)
        grid = torch.empty(a, 2, dtype=torch.float32, device=device)
        st = torch.empty(a, 1, dtype=torch.float32, device=device)
        hs = torch.tensor([[h, w] for h, w in shapes], dtype=torch.int32)
        ss = torch.tensor([float(s) for s in strides], dtype=torch.float32)
        with torch.cuda.device(device):
            rc = _cabi.lib().yb_make_anchors(_cabi.ctypes.c_void_p(hs.data_ptr()), _cabi.ctypes.c_void_p(ss.data_ptr()),
                                             len(shapes), _cabi.ptr(grid), _cabi.ptr(st), _cabi.stream_ptr(device))
        _cabi.check(rc, "yb_make_anchors")
        # integer cell indices -> dtype, then + offset in dtype: the reference's rounding sequence
        hit = (grid.to(dtype) + offset, st.to(dtype))
        _anchor_cache[key] = hit
    return hit[0].clone(), hit[1].clone()


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """ltrb distances -> xywh / xyxy boxes.  The CUDA kernel covers the layout the reference uses
    (``distance (N, 4, A)``, ``anchor_points (1, 2, A)`` or ``(2, A)``, ``dim=1``,
    src/model/model_builder.py:130); other layouts are permuted into it."""
    _cabi.require_cuda(distance, "distance")
    d = distance
    nd = d.dim()
    dim = dim % nd
    if d.shape[dim] != 4:
        raise ValueError(f"dist2bbox: size of dim {dim} must be 4, got {d.shape[dim]}")
    # bring to (N, 4, A)
    moved = d.movedim(dim, -2) if dim != nd - 2 else d
    lead = moved.shape[:-2]
    a = moved.shape[-1]
    ltrb = moved.reshape(-1, 4, a).float().contiguous()
    anc = anchor_points
    if anc.dim() == nd:
        anc = anc.movedim(dim, -2) if dim != nd - 2 else anc
        if anc.numel() != 2 * a:
            raise ValueError("dist2bbox: anchor_points must broadcast over the batch (one grid for all images)")
        anc = anc.reshape(2, a)
    elif anc.shape != (2, a):
        raise ValueError(f"dist2bbox: anchor_points must be (2, {a}) or (1, 2, {a}) for dim=1, got {tuple(anc.shape)}")
    anc = anc.to(device=d.device, dtype=torch.float32).contiguous()
    out = torch.empty_like(ltrb)
    if ltrb.numel():
        with torch.cuda.device(d.device):
            rc = _cabi.lib().yb_dist2bbox(_cabi.ptr(ltrb), _cabi.ptr(anc), ltrb.shape[0], a, int(bool(xywh)),
                                          _cabi.ptr(out), _cabi.stream_ptr(d.device))
        _cabi.check(rc, "yb_dist2bbox")
    out = out.reshape(*lead, 4, a).to(distance.dtype)
    return out.movedim(-2, dim) if dim != nd - 2 else out


def box_iou(box1, box2, eps=1e-7):
    """Pairwise IoU of xyxy boxes ``(N, 4)`` x ``(M, 4)`` -> ``(N, M)`` (reference :131-151)."""
    _cabi.require_cuda(box1, "box1")
    _cabi.require_cuda(box2, "box2")
    b1 = box1.detach().float().contiguous()
    b2 = box2.detach().float().contiguous()
    n, m = b1.shape[0], b2.shape[0]
    out = torch.empty(n, m, dtype=torch.float32, device=b1.device)
    for lo in range(0, n, 65535):
        hi = min(n, lo + 65535)
        if m == 0:
            break
        with torch.cuda.device(b1.device):
            rc = _cabi.lib().yb_box_iou(_cabi.ptr(b1[lo:hi]), hi - lo, _cabi.ptr(b2), m, float(eps), _cabi.ptr(out[lo:hi]),
                                        _cabi.stream_ptr(b1.device))
        _cabi.check(rc, "yb_box_iou")
    return out.to(box1.dtype)


def xywh2xyxy(x):
    """(…, 4) centre-size boxes -> corner boxes (reference :153-172)."""
    assert x.shape[-1] == 4, f'input shape last dimension expected 4 but input shape is {x.shape}'
    _cabi.require_cuda(x, "x")
    src = x.detach().float().contiguous()
    out = torch.empty_like(src)
    n = src.numel() // 4
    if n:
        with torch.cuda.device(x.device):
            rc = _cabi.lib().yb_xywh2xyxy(_cabi.ptr(src), n, _cabi.ptr(out), _cabi.stream_ptr(x.device))
        _cabi.check(rc, "yb_xywh2xyxy")
    return out.to(x.dtype)


def batched_nms_raw(prediction: torch.Tensor, conf_thres: float, iou_thres: float, max_det: int, nc: int,
                    agnostic: bool = False, classes=None, want_anchor: bool = False, multi_label: bool = False):
    """One ``yb_nms`` (or, with ``multi_label``, ``yb_nms_multilabel``) call.  Returns device tensors
    ``(rows (N, max_det, 6), count (N,), anchor or None)`` without any host synchronisation."""
    _cabi.require_cuda(prediction, "prediction")
    pred = prediction.detach()
    if pred.dtype != torch.float32:
        pred = pred.float()
    pred = pred.contiguous()
    n, ch, a = pred.shape
    if ch != 4 + nc:
        # the reference fails at `x.split((4, nc), 1)` when mask channels are present (:238)
        raise RuntimeError(f"split_with_sizes expects split_sizes to sum exactly to {ch} "
                           f"(input tensor's size at dimension 1), but got split_sizes=[4, {nc}]")
    dev = pred.device
    lib = _cabi.lib()
    ws_bytes = lib.yb_nms_multilabel_workspace_bytes(n) if multi_label else lib.yb_nms_workspace_bytes(n, a)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    rows = torch.empty(n, max_det, 6, dtype=torch.float32, device=dev)
    count = torch.empty(n, dtype=torch.int32, device=dev)
    anchor = torch.empty(n, max_det, dtype=torch.int32, device=dev) if want_anchor else None
    filt = None
    if classes is not None:
        filt = torch.tensor([int(c) for c in classes], dtype=torch.int32).to(dev)
    entry = lib.yb_nms_multilabel if multi_label else lib.yb_nms
    with torch.cuda.device(dev):
        rc = entry(_cabi.ptr(pred), n, nc, a, float(conf_thres), float(iou_thres), int(max_det), int(bool(agnostic)),
                   _cabi.ptr(filt), 0 if filt is None else filt.numel(), _cabi.ptr(rows), _cabi.ptr(count),
                   _cabi.ptr(anchor), _cabi.ptr(ws), ws.numel(), _cabi.stream_ptr(dev))
    _cabi.check(rc, "yb_nms_multilabel" if multi_label else "yb_nms")
    return rows, count, anchor


def non_max_suppression(
        prediction,
        conf_thres=0.25,
        iou_thres=0.45,
        classes=None,
        agnostic=False,
        multi_label=False,
        labels=(),
        max_det=300,
        nc=0,  # number of classes (optional)
):
    """Batched class-aware NMS with the reference's signature and output format.

    Returns a list with one ``(n, 6)`` tensor ``[x1, y1, x2, y2, conf, cls]`` per image, score
    descending, on ``prediction.device`` (reference :174-279).  All images are processed by three
    kernel launches and ONE device-to-host copy (the per-image counts); the reference's wall-clock
    abort (:212, :275-277) is deliberately not reproduced — every image is processed.
    ``multi_label=True`` (with nc > 1, :213) makes every (anchor, class) pair above ``conf_thres`` a candidate
    (:240-242) and keeps the 30000 best of an image (:259): six launches, no (A x nc) candidate list.
    Non-empty apriori ``labels``
    raise the RuntimeError the reference raises: it builds label rows ``nc + nm + 5`` wide (:227, the
    YOLOv5 layout with an objectness column) and cannot concatenate them with its ``4 + nc + nm`` wide
    candidates (:231).
    """
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    multi_label = bool(multi_label) and nc > 1             # :213
    if labels and any(len(lb) for lb in labels):
        width = prediction.shape[1]
        raise RuntimeError(f"Sizes of tensors must match except in dimension 0. Expected size {width} but got size "
                           f"{width + 1} for tensor number 1 in the list.")
    if classes is not None and len(classes) == 0:
        return [torch.zeros((0, 6), device=prediction.device)] * bs
    rows, count, _ = batched_nms_raw(prediction, conf_thres, iou_thres, max_det, nc, agnostic, classes,
                                     multi_label=multi_label)
    counts = count.tolist()                                  # the only host sync
    empty = torch.zeros((0, 6), device=prediction.device)
    return [rows[i, :c] if c else empty for i, c in enumerate(counts)]
