"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; this is what pins every later parity claim."""
import numpy as np
import pytest
import torch
import torchvision

from conftest import golden_loss_inputs, load_golden, unpack_gt
from oracle import decode_oracle as D
from oracle import loss_oracle as L
from oracle import nms_oracle as N
from custom_yolo_implmentation_b200.utils import synthetic as syn

LOSS_RTOL = 1e-5     # north_star: fp32 loss values and gradients within 1e-5 relative
BF16_RTOL = 1e-2     # north_star: bf16 inputs within 1e-2 relative


def _grad_close(got, ref, rtol):
    ref = ref.float(); got = got.float()
    scale = ref.abs().max().item()
    return (got - ref).abs().max().item() <= rtol * scale


@pytest.mark.parametrize("name", ["loss_small_fp32", "loss_conflict_fp32", "loss_nc171_fp32", "loss_small_bf16"])
@pytest.mark.parametrize("impl", ["spec", "torch"])
def test_loss_oracle_matches_reference(name, impl):
    z, preds, gts, anchors, strides, grad = golden_loss_inputs(name)
    nc = int(z["meta"][1])
    tr = L.loss_forward_backward(preds, gts, anchors, strides, nc, match_impl=impl)
    rtol = BF16_RTOL if z["meta"][5] else LOSS_RTOL
    assert abs(tr.total.item() - float(z["total_loss"])) <= rtol * abs(float(z["total_loss"]))
    assert abs(tr.dfl_mean.item() - float(z["box_loss"])) <= rtol * abs(float(z["box_loss"]))
    assert abs(tr.cls_mean.item() - float(z["cls_loss"])) <= rtol * abs(float(z["cls_loss"]))
    for b, idx in enumerate(tr.idx):            # matched anchors: bit-exact
        m = int(z["gt_count"][b])
        assert idx.tolist() == z["idx"][b, :m].tolist()
        assert torch.allclose(tr.iou[b], torch.from_numpy(z["iou"][b, :m]), rtol=1e-6, atol=1e-7)
    assert tr.grad.dtype == preds.dtype
    assert _grad_close(tr.grad, grad, rtol)


def test_conflict_fixture_really_has_duplicates():
    z = load_golden("loss_conflict_fp32")
    dup = sum(int(z["gt_count"][b]) - len(set(z["idx"][b, : int(z["gt_count"][b])].tolist())) for b in range(z["idx"].shape[0]))
    assert dup >= 3


def test_cfg1_summary():
    """cfg1 (N=16, 640x640, nc=80, <=50 GT) — inputs regenerated from the seed, outputs from the reference."""
    z = load_golden("loss_cfg1_summary")
    n, nc, imgsz, gmax, seed = (int(v) for v in z["meta"][:5])
    preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, seed)
    assert [g.shape[0] for g in gts] == z["gt_count"].tolist()
    tr = L.loss_forward_backward(preds, gts, anchors, strides, nc)
    assert abs(tr.total.item() - float(z["total_loss"])) <= LOSS_RTOL * float(z["total_loss"])
    agree = sum(int((tr.idx[b].numpy() == z["idx"][b, : len(tr.idx[b])]).sum()) for b in range(n))
    assert agree == int(z["gt_count"].sum())
    g = tr.grad.flatten()
    samp = g[:: int(z["grad_sample_stride"])]
    assert (samp - torch.from_numpy(z["grad_sample"])).abs().max().item() <= LOSS_RTOL * float(z["grad_absmax"])
    assert abs(g.double().abs().sum().item() - float(z["grad_abs_sum"])) <= 1e-5 * float(z["grad_abs_sum"])


def test_spec_distance_is_cdist_before_the_sqrt():
    """The stated K=4 FMA order reproduces ATen's matmul-form squared distance bit for bit (M >= 2)."""
    g = torch.Generator().manual_seed(5)
    for m in (2, 3, 7, 50, 100):
        gt = torch.rand(m, 2, generator=g) * 640
        pr = torch.rand(8400, 2, generator=g) * 700 - 30
        spec = L.center_distance(gt, pr, "spec")
        ref = torch.cdist(gt, pr)
        # the vectorised CPU sqrt of this torch build is 1 ulp off on ~0.5 % of inputs; never more
        ulp = torch.abs(spec.view(torch.int32) - ref.view(torch.int32))
        assert int(ulp.max()) <= 1
        assert (ulp != 0).float().mean().item() < 0.02
        assert (spec.argmin(1) == ref.argmin(1)).float().mean().item() >= 0.99


def test_spec_distance_single_gt_row():
    """M == 1: MKL routes the (1x4)x(4xA) product to a gemv whose accumulation order is not the
    GEMM one, so only closeness (not bit equality) holds against torch.cdist on this machine."""
    g = torch.Generator().manual_seed(6)
    gt = torch.rand(1, 2, generator=g) * 640
    pr = torch.rand(8400, 2, generator=g) * 700 - 30
    spec, ref = L.center_distance(gt, pr, "spec"), torch.cdist(gt, pr)
    assert torch.allclose(spec, ref, rtol=1e-4, atol=2e-2)
    assert spec.argmin(1).item() == ref.argmin(1).item()


def test_all_empty_batch_raises_like_the_reference():
    preds, _, anchors, strides = syn.make_loss_inputs(2, 4, 64, 3, 1)
    with pytest.raises(AttributeError):
        L.loss_forward(preds, [torch.zeros(0, 5), torch.zeros(0, 5)], anchors, strides, 4)


@pytest.mark.parametrize("name", ["nms_small", "nms_dense_maxdet", "nms_agnostic", "nms_classes", "nms_few"])
def test_nms_oracle_matches_reference(name):
    z = load_golden(name)
    n, nc, _, _, max_det, agnostic, _ = (int(v) for v in z["meta"])
    classes = z["classes"].tolist() or None
    tr = N.nms_forward(torch.from_numpy(z["prediction"]), float(z["conf"]), float(z["iou"]), classes=classes,
                       agnostic=bool(agnostic), max_det=max_det, nc=nc)
    for b in range(n):
        k = int(z["count"][b])
        assert tr.rows[b].shape[0] == k
        assert np.array_equal(tr.rows[b].numpy(), z["rows"][b, :k])        # bit-exact rows


def test_c_greedy_nms_is_torchvision_nms():
    g = torch.Generator().manual_seed(9)
    for n in (1, 2, 63, 64, 65, 1000, 3000):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = 4 + torch.rand(n, 2, generator=g) * 60
        boxes = torch.cat((xy, xy + wh), 1)
        scores = torch.rand(n, generator=g)
        order = scores.argsort(descending=True, stable=True)
        for thr in (0.0, 0.3, 0.7, 1.0):
            ref = torchvision.ops.nms(boxes, scores, thr)
            got = order[N.nms_greedy_sorted(boxes[order], thr, n)]
            assert torch.equal(got, ref)


@pytest.mark.parametrize("name", ["decode_topk", "decode_sparse"])
def test_decode_oracle_matches_reference(name):
    z = load_golden(name)
    n, nc, _, _, top_k = (int(v) for v in z["meta"])
    preds, anchors, strides = (torch.from_numpy(z[k]) for k in ("preds", "anchors", "strides"))
    tr = D.val_decode(preds, anchors, strides, float(z["conf"]), top_k, nc)
    for b in range(n):
        k = int(z["count"][b])
        assert tr.rows[b].shape[0] == k
        assert torch.allclose(tr.rows[b], torch.from_numpy(z["rows"][b, :k]), rtol=1e-6, atol=1e-5)
        assert torch.equal(tr.rows[b][:, 4], torch.from_numpy(z["rows"][b, :k, 4]))      # class ids exact
    ltrb = D.dfl_expectation(preds[:, :64, :])
    assert torch.allclose(ltrb, torch.from_numpy(z["dfl_ltrb"]), rtol=1e-6, atol=1e-6)
    box = D.ltrb_to_box(ltrb, anchors.unsqueeze(0), xywh=True, dim=1) * strides
    assert torch.allclose(box, torch.from_numpy(z["box_xywh"]), rtol=1e-5, atol=1e-4)   # DFL is a 1x1 conv there
    xyxy = D.ltrb_to_box(ltrb, anchors.unsqueeze(0), xywh=False, dim=1)
    assert torch.allclose(xyxy, torch.from_numpy(z["box_xyxy_grid"]), rtol=1e-6, atol=1e-5)


def test_helper_oracles_match_reference():
    z = load_golden("helpers")
    t = lambda k: torch.from_numpy(z[k])
    assert torch.allclose(L.iou_xywh_reference(t("b1"), t("b2")), t("bbox_iou"), rtol=1e-6, atol=1e-7)
    assert torch.allclose(L.qfl_sum(t("scores"), t("target")), t("qfl"), rtol=1e-6)
    assert torch.allclose(L.dfl_loss_rows(t("dist"), t("tval")).mean(), t("dfl"), rtol=1e-6)
    assert torch.equal(N.xywh_to_xyxy(t("b1")), t("xyxy"))
    assert torch.allclose(D.pairwise_iou_xyxy(t("xyxy"), N.xywh_to_xyxy(t("q"))), t("box_iou"), rtol=1e-6, atol=1e-7)
    assert torch.allclose(D.pairwise_iou_xywh(t("b1"), t("q")), t("box_iou_batch"), rtol=1e-6, atol=1e-7)
    anc, st = D.make_anchor_grid([(6, 5), (3, 3), (2, 1)], [8, 16, 32])
    assert torch.equal(anc, t("anchors")) and torch.equal(st, t("strides"))
    a2, s2 = syn.anchor_grid(64)
    ar, sr = D.make_anchor_grid([(8, 8), (4, 4), (2, 2)], [8, 16, 32])
    assert torch.equal(a2, ar.t()) and torch.equal(s2, sr.t())


@pytest.mark.parametrize("name", ["head_aligned", "head_ragged"])
def test_head_tail_oracle_matches_reference(name):
    z = load_golden(name)
    n, nc, nl = (int(v) for v in z["meta"])
    box = [torch.from_numpy(z[f"box{i}"]) for i in range(nl)]
    cls = [torch.from_numpy(z[f"cls{i}"]) for i in range(nl)]
    x = D.head_tail(box, cls)
    assert x.shape == (n, 64 + nc, sum(int(h) * int(w) for h, w in z["shapes"]))
    assert np.array_equal(x.numpy(), z["x"])                                   # a copy: bit-exact
    grid, st = D.make_anchor_grid([tuple(int(v) for v in s) for s in z["shapes"]], [8.0, 16.0, 32.0][:nl])
    assert np.array_equal(grid.t().numpy(), z["anchors"]) and np.array_equal(st.t().numpy(), z["strides"])


def _multilabel_prediction(z):
    n, nc, imgsz, seed = (int(v) for v in z["meta"][:4])
    if z["prediction"].size:
        return torch.from_numpy(z["prediction"])
    return syn.make_nms_input(n, nc, imgsz, seed)                              # inputs regenerated from the seed


@pytest.mark.parametrize("name", ["nms_multilabel", "nms_multilabel_classes", "nms_multilabel_cap"])
def test_nms_oracle_multi_label_matches_reference(name):
    """multi_label=True (model_utils.py:240-242); the _cap case has 672 000 candidates, cut to max_nms = 30 000."""
    z = load_golden(name)
    n, nc, _, _, max_det, agnostic, _ = (int(v) for v in z["meta"])
    classes = z["classes"].tolist() or None
    tr = N.nms_forward(_multilabel_prediction(z), float(z["conf"]), float(z["iou"]), classes=classes, agnostic=bool(agnostic),
                       max_det=max_det, nc=nc, multi_label=True)
    for b in range(n):
        k = int(z["count"][b])
        assert tr.rows[b].shape[0] == k
        assert np.array_equal(tr.rows[b].numpy(), z["rows"][b, :k])
