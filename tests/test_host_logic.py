"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the host
logic (GT packing, argument checking, sharding) behaves, and nothing silently falls back to the CPU."""
import os
import re

import pytest
import torch

from conftest import ROOT
from custom_yolo_implmentation_b200 import _cabi
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200.training import distributed_setup as DS
from custom_yolo_implmentation_b200.training.train_model import decode_predictions
from custom_yolo_implmentation_b200.utils import model_utils as U
from custom_yolo_implmentation_b200.utils import synthetic as syn


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "yolo_boxpath.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.lib()                      # raises if the extension was not built
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/yolo_boxpath.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared, "the ctypes binding and the header disagree"
    assert lib.yb_abi_version() == _cabi.ABI_VERSION
    # size queries are pure host code and safe without a GPU
    assert lib.yb_loss_workspace_bytes(128, 8400, 6400, 0) > 6400 * 12
    assert lib.yb_nms_workspace_bytes(64, 8400) >= 64 * 16384 * 8
    assert lib.yb_loss_workspace_bytes(0, 8400, 0, 0) == 0


def test_ctypes_struct_layouts_match_the_library():
    """The host structs cross the boundary by pointer: the ctypes mirror must have the library's sizes (a stale or
    re-ordered field would corrupt arguments silently) and the hand-packed GT table its 24-byte entries."""
    import ctypes
    lib = _cabi.lib()
    assert lib.yb_struct_size(0) == ctypes.sizeof(_cabi.TalParams)
    assert lib.yb_struct_size(1) == ctypes.sizeof(_cabi.TalGrid)
    assert lib.yb_struct_size(2) == ctypes.sizeof(_cabi.PeerExchangeStruct)
    assert lib.yb_struct_size(3) == 24                  # model/losses.py::_pack_gt_gather writes these by hand
    assert lib.yb_struct_size(99) == 0
    assert _cabi.TalParams.flags.offset == ctypes.sizeof(_cabi.TalParams) - 4


def test_argument_errors_come_back_through_the_abi():
    lib = _cabi.lib()
    rc = lib.yb_loss_fwd_bwd(None, 0, 1, 1, 16, 4, None, None, None, None, 0, 0, 1.0, 1.5, None, None, None, None, None,
                             None, 0, 0, None, None, None)
    assert rc == -1 and b"null pointer" in lib.yb_last_error()
    rc = lib.yb_nms(None, 1, 1, 4, 0.1, 0.5, 300, 0, None, 0, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(RuntimeError, match="yb_nms failed"):
        _cabi.check(rc, "yb_nms")


def test_no_cpu_fallback():
    preds, gts, anchors, strides = syn.make_loss_inputs(2, 4, 64, 3, 1)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        P.YoloDFLQFLoss(num_classes=4)(preds, gts, anchors, strides)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        U.non_max_suppression(syn.make_nms_input(1, 4, 64, 1), nc=4)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        decode_predictions(preds, anchors, strides, num_classes=4)
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        P.bbox_iou(torch.zeros(2, 4), torch.zeros(2, 4))


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libyolo_boxpath.so")
    with pytest.raises(_cabi.ExtensionMissing, match="no CPU fallback"):
        _cabi.lib()


def test_pack_gt_layout():
    gts = [torch.arange(10.).view(2, 5), torch.zeros(0, 5), torch.ones(3, 5, dtype=torch.float64), torch.zeros(0)]
    gt, off, counts = P.pack_gt(gts, "cpu")
    assert counts == [2, 0, 3, 0] and off.tolist() == [0, 2, 2, 5, 5] and off.dtype == torch.int32
    assert gt.shape == (5, 5) and gt.dtype == torch.float32 and gt.is_contiguous()
    assert torch.equal(gt[:2], gts[0]) and torch.equal(gt[2:], torch.ones(3, 5))
    gt0, off0, c0 = P.pack_gt([torch.zeros(0, 5)] * 2, "cpu")
    assert gt0.shape == (0, 5) and off0.tolist() == [0, 0, 0] and c0 == [0, 0]
    with pytest.raises(ValueError):
        P.pack_gt([torch.zeros(3, 4)], "cpu")
    with pytest.raises(TypeError):
        P.pack_gt([[1, 2, 3, 4, 5]], "cpu")


def test_loss_module_contract():
    crit = P.YoloDFLQFLoss()
    assert (crit.num_classes, crit.lambda_box, crit.lambda_cls, crit.lambda_dfl, crit.reg_max) == (171, 1.5, 1.0, 1.5, 16)
    assert list(crit.parameters()) == [] and list(crit.buffers()) == []      # stateless, DDP/FSDP safe
    with pytest.raises(IndexError):
        crit(torch.zeros(2, 64 + 171, 8), [torch.zeros(0, 5)], torch.zeros(2, 8), torch.zeros(1, 8))


def test_nms_argument_checks_happen_before_any_device_work():
    x = syn.make_nms_input(1, 4, 64, 1)
    with pytest.raises(AssertionError, match="Invalid Confidence"):
        U.non_max_suppression(x, conf_thres=-0.1)
    with pytest.raises(AssertionError, match="Invalid IoU"):
        U.non_max_suppression(x, iou_thres=1.1)
    with pytest.raises(RuntimeError, match="CUDA"):                              # multi_label is implemented: it reaches the device check
        U.non_max_suppression(x, multi_label=True, nc=4)
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):       # the reference's own failure (:227-231)
        U.non_max_suppression(x, labels=[torch.tensor([[1.0, 10.0, 10.0, 5.0, 5.0]])] * x.shape[0], nc=4)


def test_synthetic_generators_are_seeded_and_shaped():
    a, s = syn.anchor_grid(640)
    assert a.shape == (2, 8400) and s.shape == (1, 8400) and syn.num_anchors(1280) == 33600
    assert a[:, 0].tolist() == [0.5, 0.5] and s[0, -1].item() == 32 and a[:, 6400].tolist() == [0.5, 0.5]
    p1, g1, _, _ = syn.make_loss_inputs(4, 80, 640, 50, 7)
    p2, g2, _, _ = syn.make_loss_inputs(4, 80, 640, 50, 7)
    assert torch.equal(p1, p2) and all(torch.equal(x, y) for x, y in zip(g1, g2))
    assert g1[0].shape[0] == 50 and g1[1].shape[0] == 0 and p1.shape == (4, 144, 8400)
    gc = syn.make_gt(4, 80, 640, 50, 9, conflict_frac=0.1)
    assert all(g.shape[0] <= 50 for g in gc)
    x = syn.make_nms_input(2, 80, 640, 3)
    assert x.shape == (2, 84, 8400) and float(x[:, 4:].min()) > 0 and float(x[:, 4:].max()) < 1


def test_shard_batch():
    assert [DS.shard_batch(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    assert DS.reduce_value(3.5) == 3.5                  # not distributed: identity


def test_packed_gt_host_and_collate():
    from custom_yolo_implmentation_b200.data.collate import collate_fn, collate_fn_packed
    gts = [torch.arange(10.).view(2, 5), torch.zeros(0, 5), torch.ones(3, 5)]
    pk = P.pack_gt_host(gts, pin_memory=False)
    assert len(pk) == 3 and pk.counts == [2, 0, 3] and pk.offsets.tolist() == [0, 2, 2, 5] and pk.gt.shape == (5, 5)
    assert torch.equal(pk.gt[2:], torch.ones(3, 5))
    batch = [(torch.zeros(3, 4, 4), {"boxes": g}) for g in gts]
    images, targets, packed = collate_fn_packed(batch)
    assert images.shape == (3, 3, 4, 4) and len(targets) == 3 and packed.counts == [2, 0, 3]
    images2, targets2 = collate_fn(batch)
    assert torch.equal(images, images2) and targets2[2]["boxes"] is gts[2]


def test_grid_hint_is_pure_host_logic():
    """build_grid_hint describes the reference's pyramid of grids (make_anchors, model_utils.py:60-70) or returns None;
    exact=False also takes anchors built in bf16 (159.5 rounds to 160), which only the nearest-centre path's
    probe role may use -- it needs a cell NEAR each box centre, nothing more."""
    from custom_yolo_implmentation_b200.model.losses import build_grid_hint
    anchors, strides = syn.anchor_grid(1280)
    hint = build_grid_hint(anchors, strides)
    assert hint is not None and hint.n_levels == 3
    assert list(hint.w)[:3] == [160, 80, 40] and list(hint.start)[:3] == [0, 25600, 32000]
    assert list(hint.stride)[:3] == [8.0, 16.0, 32.0] and hint.x0[0] == 0.5 and hint.y0[2] == 0.5
    ab, sb = anchors.bfloat16().float(), strides.bfloat16().float()
    assert build_grid_hint(ab, sb) is None                          # not bit-for-bit a grid any more
    near = build_grid_hint(ab, sb, exact=False)
    assert near is not None and list(near.w)[:3] == [160, 80, 40]
    perm = torch.randperm(anchors.shape[1], generator=torch.Generator().manual_seed(0))
    assert build_grid_hint(anchors[:, perm], strides[:, perm]) is None
    assert build_grid_hint(anchors[:, perm], strides[:, perm], exact=False) is None
    assert build_grid_hint(anchors[:, :7], strides[:, :9]) is None  # mismatched lengths
