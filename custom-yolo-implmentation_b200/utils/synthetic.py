"""Seeded synthetic inputs for the box-geometry hot path.

Every tensor is drawn from a CPU ``torch.Generator`` so the CPU oracle and the CUDA path
see identical bits; callers move the result to the device they want.  Nothing here is on the
product path: it feeds ``bench.py``, ``__graft_entry__.smoke()`` and ``tests/``.

Shapes follow the reference's head output contract (``src/model/head.py:112-121``):
``preds (N, 4*reg_max + nc, A)`` channel-major with the anchor axis contiguous,
``anchors (2, A)`` in grid units, ``strides (1, A)``, and the data layer's per-image list of
``(Mi, 5)`` float32 ``[cx, cy, w, h, cls]`` pixel boxes (``src/data/dataset_loader.py:76``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

DEFAULT_STRIDES = (8, 16, 32)


def grid_shapes(imgsz: int, strides: Sequence[int] = DEFAULT_STRIDES) -> List[Tuple[int, int]]:
    """Feature-map (h, w) per pyramid level for a square ``imgsz`` input."""
    return [(imgsz // s, imgsz // s) for s in strides]


def num_anchors(imgsz: int, strides: Sequence[int] = DEFAULT_STRIDES) -> int:
    return sum(h * w for h, w in grid_shapes(imgsz, strides))


def anchor_grid(imgsz: int, strides: Sequence[int] = DEFAULT_STRIDES, offset: float = 0.5,
                dtype: torch.dtype = torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Cell-centre grid, x fastest, P3->P4->P5, returned as ``(2, A)`` / ``(1, A)``.

    Same values as the reference's ``make_anchors`` (``src/utils/model_utils.py:60-70``)
    followed by the transposes of ``Head.forward`` (``src/model/head.py:112-114``).
    Built on the CPU with index arithmetic only (exact halves), then cast to ``dtype``.
    """
    ax, ay, st = [], [], []
    for (h, w), s in zip(grid_shapes(imgsz, strides), strides):
        col = (torch.arange(w, dtype=torch.float32) + offset).to(dtype)
        row = (torch.arange(h, dtype=torch.float32) + offset).to(dtype)
        ax.append(col.repeat(h))
        ay.append(row.repeat_interleave(w))
        st.append(torch.full((h * w,), float(s), dtype=dtype))
    anchors = torch.stack((torch.cat(ax), torch.cat(ay)), 0).contiguous()
    return anchors, torch.cat(st).unsqueeze(0).contiguous()


def make_preds(n: int, nc: int, a: int, seed: int, reg_max: int = 16,
               cls_mean: float = -4.595, cls_std: float = 1.0,
               dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Head output: box logits ~ N(0,1), class logits ~ N(log(0.01/0.99), 1).

    The class-logit mean is the head's bias initialisation (``src/model/head.py:68-69``).
    Generated in fp32 and cast, so a bf16 run sees the rounded fp32 draw.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    c = 4 * reg_max + nc
    x = torch.randn(n, c, a, generator=g, dtype=torch.float32)
    x[:, 4 * reg_max:, :].mul_(cls_std).add_(cls_mean)
    return x.to(dtype).contiguous()


def make_gt(n: int, nc: int, imgsz: int, gmax: int, seed: int, conflict_frac: float = 0.0,
            force_empty_and_full: bool = True) -> List[torch.Tensor]:
    """Per-image GT lists ``(Mi, 5)`` with ``Mi ~ U{0..gmax}``.

    With ``force_empty_and_full`` image 0 gets ``gmax`` boxes and (when n > 1) image 1 gets
    none; the batch is never all-empty (the reference raises on that).
    ``conflict_frac`` duplicates that share of each image's boxes with a +0.5 px jitter so two
    GTs land on one anchor.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = []
    for i in range(n):
        m = int(torch.randint(0, gmax + 1, (1,), generator=g).item())
        if force_empty_and_full:
            if i == 0:
                m = gmax
            elif i == 1:
                m = 0
        cxcy = torch.rand(m, 2, generator=g) * imgsz
        wh = 8.0 + torch.rand(m, 2, generator=g) * (0.4 * imgsz - 8.0)
        cls = torch.randint(0, nc, (m, 1), generator=g).float()
        boxes = torch.cat((cxcy, wh, cls), 1)
        if conflict_frac > 0 and m > 1:
            k = min(max(1, int(round(conflict_frac * m))), m // 2)
            keep = min(m, gmax - k)  # originals that survive the gmax cap
            src = torch.randperm(keep, generator=g)[:k]
            dup = boxes[src].clone()
            dup[:, :2] += 0.5
            dup[:, 4:5] = torch.randint(0, nc, (dup.shape[0], 1), generator=g).float()
            boxes = torch.cat((boxes[:keep], dup), 0)
        out.append(boxes.contiguous())
    return out


def make_loss_inputs(n: int, nc: int, imgsz: int, gmax: int, seed: int,
                     dtype: torch.dtype = torch.float32, conflict_frac: float = 0.0):
    """(preds, gt_list, anchors, strides) for the loss path, all on the CPU."""
    anchors, strides = anchor_grid(imgsz, dtype=dtype)
    a = anchors.shape[1]
    preds = make_preds(n, nc, a, seed, dtype=dtype)
    gts = make_gt(n, nc, imgsz, gmax, seed + 7919, conflict_frac=conflict_frac)
    return preds, gts, anchors, strides


def make_nms_input(n: int, nc: int, imgsz: int, seed: int, logit_mean: float = -3.0,
                   logit_std: float = 2.0, dense_uniform: bool = False) -> torch.Tensor:
    """NMS input ``(N, 4 + nc, A)``: xywh pixel boxes + per-class scores in (0, 1).

    Boxes sit near their anchor cell with sizes of a few cells, so neighbouring candidates
    overlap and suppression does real work; scores are ``sigmoid(N(mean, std))`` (≈ all of
    them pass conf 0.001) or ``U(0,1)`` for the dense variant.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    anchors, strides = anchor_grid(imgsz)
    a = anchors.shape[1]
    ctr = (anchors * strides).unsqueeze(0) + (torch.rand(n, 2, a, generator=g) - 0.5) * strides * 2.0
    wh = strides * (1.0 + 7.0 * torch.rand(n, 2, a, generator=g))
    if dense_uniform:
        sc = torch.rand(n, nc, a, generator=g)
    else:
        sc = torch.sigmoid(torch.randn(n, nc, a, generator=g) * logit_std + logit_mean)
    return torch.cat((ctr, wh, sc), 1).contiguous()
