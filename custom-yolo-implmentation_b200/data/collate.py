"""Collate function with the kernels' GT wire format.

The reference's ``collate_fn`` (src/data/collate.py:3-9) returns ``(images, [target, ...])`` and the
training loop then copies every image's boxes to the device separately
(src/training/train_model.py:236).  ``collate_fn_packed`` keeps that tuple shape but adds the packed,
pinned GT buffer so that the loss needs two small async copies per batch.
"""
from __future__ import annotations

import torch

from ..model.losses import pack_gt_host

__all__ = ["collate_fn", "collate_fn_packed"]


def collate_fn(batch):
    """Same behaviour as the reference's ``collate_fn``."""
    images = torch.stack([item[0] for item in batch])
    targets = [item[1] for item in batch]
    return images, targets


def collate_fn_packed(batch):
    """``(images, targets, packed_gt)``; ``targets`` as in the reference (dicts with a ``"boxes"`` entry or
    plain ``(Mi, 5)`` tensors), ``packed_gt`` a pinned ``PackedGT`` for ``YoloDFLQFLoss.forward``."""
    images, targets = collate_fn(batch)
    boxes = [t["boxes"] if isinstance(t, dict) else t for t in targets]
    return images, targets, pack_gt_host(boxes)
