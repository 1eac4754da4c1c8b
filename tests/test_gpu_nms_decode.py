"""Parity of the CUDA decode / NMS / IoU-utility kernels (through the C ABI) against the reference's
golden vectors and the CPU oracle.  Needs a B200: run with ``-m gpu``.

Integer results (keep lists, class ids, anchor indices, row order) are compared bit-exactly; the NMS
output rows are compared bit-exactly too (they are copies / exact fp32 arithmetic of the input)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import decode_oracle as D
from oracle import loss_oracle as L
from oracle import nms_oracle as N
from custom_yolo_implmentation_b200.model import losses as PL
from custom_yolo_implmentation_b200.model.model_blocks import DFL, dfl_decode
from custom_yolo_implmentation_b200.training.metrics import box_iou_batch
from custom_yolo_implmentation_b200.training.train_model import decode_predictions, decode_predictions_raw
from custom_yolo_implmentation_b200.utils import model_utils as U
from custom_yolo_implmentation_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------ NMS
@pytest.mark.parametrize("name", ["nms_small", "nms_dense_maxdet", "nms_agnostic", "nms_classes", "nms_few"])
def test_nms_matches_reference_golden(name, cuda_device):
    z = load_golden(name)
    n, nc, _, _, max_det, agnostic, _ = (int(v) for v in z["meta"])
    classes = z["classes"].tolist() or None
    out = U.non_max_suppression(torch.from_numpy(z["prediction"]).to(cuda_device), float(z["conf"]), float(z["iou"]),
                                classes=classes, agnostic=bool(agnostic), max_det=max_det, nc=nc)
    assert len(out) == n
    for b in range(n):
        k = int(z["count"][b])
        assert out[b].shape == (k, 6)
        assert np.array_equal(out[b].cpu().numpy(), z["rows"][b, :k])          # bit-exact rows, same order


@pytest.mark.parametrize("name", ["nms_multilabel", "nms_multilabel_classes", "nms_multilabel_cap"])
def test_nms_multi_label_matches_reference_golden(name, cuda_device):
    z = load_golden(name)
    n, nc, imgsz, seed, max_det, agnostic, _ = (int(v) for v in z["meta"])
    classes = z["classes"].tolist() or None
    x = torch.from_numpy(z["prediction"]) if z["prediction"].size else syn.make_nms_input(n, nc, imgsz, seed)
    out = U.non_max_suppression(x.to(cuda_device), float(z["conf"]), float(z["iou"]), classes=classes,
                                agnostic=bool(agnostic), multi_label=True, max_det=max_det, nc=nc)
    for b in range(n):
        k = int(z["count"][b])
        assert out[b].shape == (k, 6)
        assert np.array_equal(out[b].cpu().numpy(), z["rows"][b, :k])


@pytest.mark.parametrize("agnostic", [False, True])
def test_nms_multi_label_against_oracle(agnostic, cuda_device):
    """More candidates than max_nms on a 320 px grid with 40 classes, several images, both offset modes."""
    x = syn.make_nms_input(3, 40, 320, 97)
    rows, count, anchor = U.batched_nms_raw(x.to(cuda_device), 0.001, 0.6, 300, 40, agnostic, None, want_anchor=True,
                                            multi_label=True)
    ora = N.nms_forward(x, 0.001, 0.6, agnostic=agnostic, max_det=300, nc=40, multi_label=True)
    assert min(ora.n_candidates) > 30000
    for b in range(3):
        k = int(count[b])
        assert k == ora.rows[b].shape[0]
        assert torch.equal(anchor[b, :k].cpu().long(), ora.keep_anchor[b])
        assert torch.equal(rows[b, :k].cpu(), ora.rows[b])


def _check_against_oracle(x, conf, iou, max_det, nc, dev, agnostic=False, classes=None):
    rows, count, anchor = U.batched_nms_raw(x.to(dev), conf, iou, max_det, nc, agnostic, classes, want_anchor=True)
    ora = N.nms_forward(x, conf, iou, classes=classes, agnostic=agnostic, max_det=max_det, nc=nc)
    count = count.cpu()
    for b in range(x.shape[0]):
        k = int(count[b])
        assert k == ora.rows[b].shape[0], f"image {b}: kept {k} vs oracle {ora.rows[b].shape[0]}"
        assert torch.equal(anchor[b, :k].cpu().long(), ora.keep_anchor[b])   # keep list, bit-exact, in order
        assert torch.equal(rows[b, :k].cpu(), ora.rows[b])
    return count


@pytest.mark.parametrize("dense", [False, True])
def test_nms_cfg4_shape_against_oracle(dense, cuda_device):
    """cfg4 shape (8400 candidates, nc=80, conf 0.001, IoU 0.7, max_det 300) on 4 images."""
    x = syn.make_nms_input(4, 80, 640, 4321 + dense, dense_uniform=dense)
    cnt = _check_against_oracle(x, 0.001, 0.7, 300, 80, cuda_device)
    assert int(cnt.min()) == 300


def test_nms_edge_cases(cuda_device):
    x = syn.make_nms_input(3, 5, 160, 77)
    x[1, 4:] = 0.0005                                   # image with no candidate at all
    x[2, 4:, 10:] = 0.0                                 # image with exactly 10 candidates
    _check_against_oracle(x, 0.001, 0.5, 300, 5, cuda_device)
    _check_against_oracle(x, 0.001, 0.5, 7, 5, cuda_device)               # max_det cap
    _check_against_oracle(x, 0.001, 0.0, 300, 5, cuda_device)             # IoU threshold 0
    _check_against_oracle(x, 0.001, 1.0, 300, 5, cuda_device)             # IoU threshold 1: nothing suppressed
    _check_against_oracle(x, 0.001, 0.6, 300, 5, cuda_device, agnostic=True)
    _check_against_oracle(x, 0.001, 0.6, 300, 5, cuda_device, classes=[0, 3])
    out = U.non_max_suppression(x.to(cuda_device), 0.001, 0.5, nc=5)
    assert out[1].shape == (0, 6)
    # scalar path: anchor count not a multiple of 4 / unaligned rows
    y = syn.make_nms_input(2, 3, 96, 78)
    assert y.shape[2] % 4 != 0
    _check_against_oracle(y, 0.01, 0.45, 300, 3, cuda_device)
    # tuple input, nc inferred, argument errors as in the reference
    out2 = U.non_max_suppression((x.to(cuda_device), None), 0.001, 0.5)
    assert all(torch.equal(a, b) for a, b in zip(out, out2))
    with pytest.raises(AssertionError):
        U.non_max_suppression(x.to(cuda_device), conf_thres=1.5)
    with pytest.raises(RuntimeError):                   # mask channels: the reference fails in split()
        U.non_max_suppression(x.to(cuda_device), 0.1, 0.5, nc=3)


def test_nms_more_candidates_than_register_columns(cuda_device):
    """1280x1280 grid (33600 anchors): > 9216 candidates -> the global-scratch sweep, and the 30000 cap."""
    x = syn.make_nms_input(2, 4, 1280, 91)
    _check_against_oracle(x, 0.001, 0.6, 300, 4, cuda_device)
    x[:, 4:] = x[:, 4:] * 0.9 + 0.05                  # every anchor is a candidate: 33600 > max_nms
    _check_against_oracle(x, 0.001, 0.6, 300, 4, cuda_device)


def test_nms_full_cfg4_properties(cuda_device):
    """cfg4 at full size (N=64): structural properties on all images, oracle on 3 of them."""
    n, nc, thr = 64, 80, 0.7
    x = syn.make_nms_input(n, nc, 640, 2024)
    rows, count, anchor = U.batched_nms_raw(x.to(cuda_device), 0.001, thr, 300, nc, want_anchor=True)
    count = count.cpu()
    assert int(count.min()) > 0
    for b in range(n):
        r = rows[b, : int(count[b])]
        assert bool((r[1:, 4] <= r[:-1, 4]).all())                         # score descending
        a = anchor[b, : int(count[b])].long()
        assert a.unique().numel() == a.numel()
        # no kept pair of one class overlaps above the threshold (IoU on the class-offset boxes)
        off = r[:, :4] + r[:, 5:6] * 7680.0
        iou = U.box_iou(off, off, eps=0.0)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= thr
    sel = [0, 31, 63]
    ora = N.nms_forward(x[sel], 0.001, thr, max_det=300, nc=nc)
    for k, b in enumerate(sel):
        assert torch.equal(anchor[b, : int(count[b])].cpu().long(), ora.keep_anchor[k])


# --------------------------------------------------------------------------------------- decode
@pytest.mark.parametrize("name", ["decode_topk", "decode_sparse"])
def test_decode_matches_reference_golden(name, cuda_device):
    z = load_golden(name)
    n, nc, _, _, top_k = (int(v) for v in z["meta"])
    dev = cuda_device
    preds, anchors, strides = (torch.from_numpy(z[k]).to(dev) for k in ("preds", "anchors", "strides"))
    out = decode_predictions(preds, anchors, strides, float(z["conf"]), top_k, nc)
    for b in range(n):
        k = int(z["count"][b])
        assert out[b].shape == (k, 5)
        ref = torch.from_numpy(z["rows"][b, :k])
        assert torch.equal(out[b][:, 4].cpu(), ref[:, 4])                 # class ids and row order exact
        assert torch.allclose(out[b][:, :4].cpu(), ref[:, :4], rtol=1e-5, atol=1e-4)
    # DFL module + dist2bbox + stride (src/model/model_builder.py:123-133)
    ltrb = DFL(16).to(dev)(preds[:, :64, :])
    assert torch.allclose(ltrb.cpu(), torch.from_numpy(z["dfl_ltrb"]), rtol=1e-5, atol=1e-5)
    box = U.dist2bbox(ltrb, anchors.unsqueeze(0), xywh=True, dim=1) * strides
    assert torch.allclose(box.cpu(), torch.from_numpy(z["box_xywh"]), rtol=1e-5, atol=1e-4)
    xyxy = U.dist2bbox(ltrb, anchors.unsqueeze(0), xywh=False, dim=1)
    assert torch.allclose(xyxy.cpu(), torch.from_numpy(z["box_xyxy_grid"]), rtol=1e-5, atol=1e-4)
    # fused decode in one launch
    _, fused = dfl_decode(preds, anchors, strides, box_format="xywh")
    assert torch.allclose(fused.cpu(), torch.from_numpy(z["box_xywh"]), rtol=1e-5, atol=1e-4)


def _check_val_decode(rows, anchor, k, preds_b, box_b, conf, top_k, tol=2e-6):
    """One image of decode_predictions against exact arithmetic, allowing ONLY what float rounding of the sigmoid can
    change: an anchor may enter / leave at the confidence threshold or at the top-k boundary, or swap places with a
    neighbour, if and only if the scores involved agree to `tol` relative.  Everything else is exact."""
    sc = preds_b[64:, :].double().sigmoid()                     # (nc, A) in fp64: the arbiter of near-ties
    best, cid = sc.max(0)
    a = anchor[:k].long()
    assert a.unique().numel() == k                              # no anchor twice
    n_cand = int((best >= conf).sum())
    near_thr = int(((best - conf).abs() <= tol * conf).sum())
    assert abs(k - min(n_cand, top_k)) <= near_thr, (k, n_cand, near_thr)
    got = best[a]
    assert (got >= conf * (1 - tol)).all()                      # nothing below the threshold
    if n_cand <= top_k - near_thr:
        assert torch.equal(a, a.sort().values)                  # no top-k: rows stay in anchor order (train_model.py:126-133)
    else:
        kth = best.sort(descending=True).values[min(top_k, n_cand) - 1]
        assert (got >= kth * (1 - tol)).all()                   # every row belongs to the top-k
        assert (got[1:] <= got[:-1] * (1 + tol)).all()          # ... in descending score order
        # equal scores (frequent with bf16 logits): lowest anchor first
        tie = (got[1:] == got[:-1])
        assert (a[1:][tie] > a[:-1][tie]).all()
    # class id: the first maximum, unless the two best classes tie to rounding
    top2 = sc[:, a].topk(2, 0).values
    clear = (top2[0] - top2[1]) > tol * top2[0]
    assert torch.equal(rows[:k, 4].long()[clear], cid[a][clear])
    assert torch.allclose(rows[:k, :4], box_b[a].float(), rtol=1e-5, atol=1e-4)


def test_val_decode_against_oracle_full_grid(cuda_device):
    anchors, strides = syn.anchor_grid(640)
    agree = {}
    for seed, conf, top_k, mean, dt in [(5, 0.25, 100, -1.0, torch.float32), (6, 0.5, 100, -6.0, torch.float32),
                                        (7, 0.25, 17, -2.0, torch.float32), (8, 0.25, 100, -1.0, torch.bfloat16),
                                        (9, 0.9, 100, -3.0, torch.bfloat16)]:
        preds = syn.make_preds(3, 80, anchors.shape[1], seed, cls_mean=mean, cls_std=1.5, dtype=dt)
        rows, count, anchor = decode_predictions_raw(preds.to(cuda_device), anchors.to(cuda_device), strides.to(cuda_device),
                                                     conf, top_k, 80, want_anchor=True)
        ora = D.val_decode(preds.float(), anchors, strides, conf, top_k, 80)
        ltrb = D.dfl_expectation(preds.float()[:, :64, :]).permute(0, 2, 1)
        box = D.ltrb_to_box(ltrb, anchors.float().transpose(0, 1).unsqueeze(0), xywh=True, dim=2) * strides.float().transpose(0, 1).unsqueeze(0)
        same = tot = 0
        for b in range(3):
            k = int(count[b])
            _check_val_decode(rows[b].cpu(), anchor[b].cpu(), k, preds[b].float(), box[b], conf, top_k)
            if k == ora.rows[b].shape[0]:
                same += int((anchor[b, :k].cpu().long() == ora.anchor[b]).sum())
            tot += ora.rows[b].shape[0]
        agree[(seed, str(dt))] = (same, tot)
    print("val decode rows in the oracle's exact position:", agree)
    # with fp32 logits score ties are rare: the CPU oracle's own float ordering is matched almost everywhere
    assert all(s >= 0.97 * t for (seed, dt), (s, t) in agree.items() if dt == "torch.float32")


def test_make_anchors_and_helpers_match_reference(cuda_device):
    z = load_golden("helpers")
    dev = cuda_device
    t = lambda k: torch.from_numpy(z[k]).to(dev)
    lv = [torch.zeros(1, 1, 6, 5, device=dev), torch.zeros(1, 1, 3, 3, device=dev), torch.zeros(1, 1, 2, 1, device=dev)]
    anc, st = U.make_anchors(lv, [8, 16, 32], 0.5)
    assert torch.equal(anc.cpu(), torch.from_numpy(z["anchors"])) and torch.equal(st.cpu(), torch.from_numpy(z["strides"]))
    ancb, _ = U.make_anchors([l.bfloat16() for l in lv], [8, 16, 32], 0.5)
    assert ancb.dtype == torch.bfloat16 and torch.equal(ancb.float().cpu(), torch.from_numpy(z["anchors"]))
    assert torch.equal(U.xywh2xyxy(t("b1")).cpu(), torch.from_numpy(z["xyxy"]))
    assert torch.allclose(U.box_iou(t("xyxy"), U.xywh2xyxy(t("q"))).cpu(), torch.from_numpy(z["box_iou"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(box_iou_batch(t("b1"), t("q")).cpu(), torch.from_numpy(z["box_iou_batch"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(PL.bbox_iou(t("b1"), t("b2")).cpu(), torch.from_numpy(z["bbox_iou"]), rtol=1e-6, atol=1e-7)
    assert torch.allclose(PL.quality_focal_loss(t("scores"), t("target")).cpu(), torch.from_numpy(z["qfl"]), rtol=1e-5)
    assert torch.allclose(PL.distribution_focal_loss(t("dist"), t("tval")).cpu(), torch.from_numpy(z["dfl"]), rtol=1e-5)


def test_helper_gradients_match_oracle_autograd(cuda_device):
    z = load_golden("helpers")
    dev = cuda_device
    # bbox_iou: d sum(w * iou) / d box1
    b1 = torch.from_numpy(z["b1"]).clone().requires_grad_(True)
    w = torch.linspace(0.5, 2.0, b1.shape[0])
    (L.iou_xywh_reference(b1, torch.from_numpy(z["b2"])) * w).sum().backward()
    g1 = torch.from_numpy(z["b1"]).to(dev).requires_grad_(True)
    (PL.bbox_iou(g1, torch.from_numpy(z["b2"]).to(dev)) * w.to(dev)).sum().backward()
    assert torch.allclose(g1.grad.cpu(), b1.grad, rtol=1e-4, atol=1e-7)
    # quality_focal_loss
    s = torch.from_numpy(z["scores"]).clone().requires_grad_(True)
    (L.qfl_sum(s, torch.from_numpy(z["target"])) * 3.0).backward()
    gs = torch.from_numpy(z["scores"]).to(dev).requires_grad_(True)
    (PL.quality_focal_loss(gs, torch.from_numpy(z["target"]).to(dev)) * 3.0).backward()
    assert torch.allclose(gs.grad.cpu(), s.grad, rtol=1e-4, atol=1e-7)
    # ... and with another focusing exponent (the reference's formula, src/model/losses.py:46-57, restated inline)
    for beta in (1.5, 3.0):
        s = torch.from_numpy(z["scores"]).clone().requires_grad_(True)
        tgt = torch.from_numpy(z["target"])
        p = s.sigmoid()
        ref = -(tgt * (1 - p).pow(beta) * torch.log(p + 1e-12) + (1 - tgt) * p.pow(beta) * torch.log(1 - p + 1e-12)).sum() / s.shape[0]
        ref.backward()
        gs = torch.from_numpy(z["scores"]).to(dev).requires_grad_(True)
        got = PL.quality_focal_loss(gs, tgt.to(dev), beta=beta)
        got.backward()
        assert abs(got.item() - ref.item()) <= 1e-5 * abs(ref.item())
        assert torch.allclose(gs.grad.cpu(), s.grad, rtol=1e-4, atol=1e-7)
    # distribution_focal_loss
    d = torch.from_numpy(z["dist"]).clone().requires_grad_(True)
    L.dfl_loss_rows(d, torch.from_numpy(z["tval"])).mean().backward()
    gd = torch.from_numpy(z["dist"]).to(dev).requires_grad_(True)
    PL.distribution_focal_loss(gd, torch.from_numpy(z["tval"]).to(dev)).backward()
    assert torch.allclose(gd.grad.cpu(), d.grad, rtol=1e-4, atol=1e-7)


def test_packed_gt_and_inference_postprocess(cuda_device):
    """§8(f).3/.4: the packed GT wire format gives the same loss; fused inference post-processing equals the
    reference's sequence DFL -> dist2bbox -> *stride -> cat -> NMS (golden decode fixture + NMS oracle)."""
    from custom_yolo_implmentation_b200.model.losses import YoloDFLQFLoss, pack_gt_host
    from custom_yolo_implmentation_b200.utils.postprocess import postprocess_inference
    dev = cuda_device
    preds, gts, anchors, strides = syn.make_loss_inputs(3, 6, 128, 10, 101)
    crit = YoloDFLQFLoss(num_classes=6)
    l1, d1 = crit(preds.to(dev), [g.to(dev) for g in gts], anchors.to(dev), strides.to(dev))
    l2, d2 = crit(preds.to(dev), pack_gt_host(gts).to(dev), anchors.to(dev), strides.to(dev))
    assert d1 == d2 and torch.equal(l1, l2)
    z = load_golden("decode_topk")
    x = torch.from_numpy(z["preds"]).to(dev)
    anc, st = torch.from_numpy(z["anchors"]).to(dev), torch.from_numpy(z["strides"]).to(dev)
    out = postprocess_inference(x, anc, st, 6, conf_thres=0.25, iou_thres=0.45, apply_sigmoid=True)
    # the decode itself is pinned to the reference above (test_decode_matches_reference_golden); here the NMS oracle
    # runs on the SAME decoded boxes so that a 1e-6 difference cannot flip an IoU decision
    _, box = dfl_decode(x, anc, st, want_ltrb=False, box_format="xywh")
    assert torch.allclose(box.cpu(), torch.from_numpy(z["box_xywh"]), rtol=1e-5, atol=1e-4)
    y = torch.cat((box.cpu(), torch.from_numpy(z["preds"])[:, 64:].sigmoid()), 1)
    ora = N.nms_forward(y, 0.25, 0.45, max_det=300, nc=6)
    for b in range(x.shape[0]):
        assert out[b].shape == ora.rows[b].shape and out[b].shape[0] > 0
        assert torch.equal(out[b][:, 5].cpu(), ora.rows[b][:, 5])
        assert torch.allclose(out[b].cpu(), ora.rows[b], rtol=1e-6, atol=1e-6)


def test_fused_postprocess_matches_reference_golden(cuda_device):
    """§8(f).4: yb_postprocess against the reference's own Model.inference tail (model_builder.py:123-139: split, DFL,
    dist2bbox, * strides, cat, NMS on RAW logits) captured in tests/golden/inference_post.npz, and bit for bit against the
    unfused sequence of this repo's own ops (decode kernel -> torch.cat -> yb_nms)."""
    from custom_yolo_implmentation_b200 import _cabi
    from custom_yolo_implmentation_b200.utils.postprocess import postprocess_inference
    dev = cuda_device
    z = load_golden("inference_post")
    x = torch.from_numpy(z["x"]).to(dev)
    anc, st = torch.from_numpy(z["anchors"]).to(dev), torch.from_numpy(z["strides"]).to(dev)
    n, nc = int(z["meta"][0]), int(z["meta"][1])
    before = _cabi.launch_count
    out = postprocess_inference(x, anc, st, nc, conf_thres=float(z["conf"]), iou_thres=float(z["iou"]))
    assert _cabi.launch_count - before == 4                 # decode, scan, class-parallel NMS, generic sweep: nothing else
    for b in range(n):
        ref = torch.from_numpy(z["rows"][b, : int(z["count"][b])])
        assert out[b].shape == ref.shape
        assert torch.equal(out[b][:, 5].cpu(), ref[:, 5]) and torch.equal(out[b][:, 4].cpu(), ref[:, 4])   # classes, raw-logit scores
        assert torch.allclose(out[b][:, :4].cpu(), ref[:, :4], rtol=1e-5, atol=1e-4)
    for dtype in (torch.float32, torch.bfloat16):
        xd = x.to(dtype)
        fused = postprocess_inference(xd, anc, st, nc, conf_thres=0.25, iou_thres=0.45)
        _, box = dfl_decode(xd, anc, st, want_ltrb=False, box_format="xywh")
        unfused = U.non_max_suppression(torch.cat((box, xd[:, 64:].float()), 1), conf_thres=0.25, iou_thres=0.45, nc=nc)
        for a, b in zip(fused, unfused):
            assert torch.equal(a, b)
        sig = postprocess_inference(xd, anc, st, nc, conf_thres=0.25, iou_thres=0.45, apply_sigmoid=True)
        unf = U.non_max_suppression(torch.cat((box, xd[:, 64:].float().sigmoid()), 1), conf_thres=0.25, iou_thres=0.45, nc=nc)
        for a, b in zip(sig, unf):
            assert a.shape == b.shape and torch.equal(a[:, 5], b[:, 5]) and torch.equal(a[:, :4], b[:, :4])
            assert torch.allclose(a[:, 4], b[:, 4], rtol=1e-6, atol=0)
    # the filters of the NMS signature
    only = postprocess_inference(x, anc, st, nc, conf_thres=0.25, iou_thres=0.45, classes=[1, 4], max_det=7)
    assert all(o.shape[0] <= 7 and set(o[:, 5].tolist()) <= {1.0, 4.0} for o in only)
    assert all(o.shape == (0, 6) for o in postprocess_inference(x, anc, st, nc, classes=[]))


# ------------------------------------------------------------------------------------------ head tail
@pytest.mark.parametrize("name", ["head_aligned", "head_ragged"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_tail_matches_reference_golden(name, dtype, cuda_device):
    from custom_yolo_implmentation_b200.model.head import head_tail
    z = load_golden(name)
    n, nc, nl = (int(v) for v in z["meta"])
    box = [torch.from_numpy(z[f"box{i}"]).to(cuda_device, dtype).requires_grad_(True) for i in range(nl)]
    cls = [torch.from_numpy(z[f"cls{i}"]).to(cuda_device, dtype).requires_grad_(True) for i in range(nl)]
    x, anchors, strides = head_tail(box, cls, torch.tensor([8.0, 16.0, 32.0][:nl]))
    want = torch.from_numpy(z["x"]).to(dtype)
    assert torch.equal(x.cpu(), want)                                           # a copy: bit-exact in either dtype
    assert anchors.shape == (2, x.shape[2]) and strides.shape == (1, x.shape[2])
    assert np.array_equal(anchors.float().cpu().numpy(), z["anchors"])
    assert np.array_equal(strides.float().cpu().numpy(), z["strides"])
    # adjoint: the backward of the two cats hands every level its slice of the gradient, bit-exactly
    g = torch.randn(x.shape, device=cuda_device).to(dtype)
    x.backward(g)
    ref_in = [t.detach().cpu().float().requires_grad_(True) for t in box + cls]
    D.head_tail(ref_in[:nl], ref_in[nl:]).backward(g.cpu().float())
    for got, ref in zip(box + cls, ref_in):
        assert torch.equal(got.grad.cpu().float(), ref.grad)


def test_head_tail_full_size_round_trip(cuda_device):
    """640 px, nc=80, N=16: gather equals torch's own cats, and scatter(gather(levels)) == levels."""
    from custom_yolo_implmentation_b200.model.head import gather_levels
    g = torch.Generator(device=cuda_device).manual_seed(7)
    shapes = [(80, 80), (40, 40), (20, 20)]
    box = [torch.randn(16, 64, h, w, device=cuda_device, generator=g).requires_grad_(True) for h, w in shapes]
    cls = [torch.randn(16, 80, h, w, device=cuda_device, generator=g).requires_grad_(True) for h, w in shapes]
    x = gather_levels(box, cls)
    assert x.shape == (16, 144, 8400)
    assert torch.equal(x, D.head_tail([b.detach() for b in box], [c.detach() for c in cls]))
    x.backward(x.detach())
    for t in box + cls:
        assert torch.equal(t.grad, t.detach())


def test_head_tail_rejects_bad_input(cuda_device):
    from custom_yolo_implmentation_b200.model.head import head_tail
    b = [torch.zeros(1, 64, 2, 2, device=cuda_device)]
    with pytest.raises(ValueError):
        head_tail(b, [torch.zeros(1, 3, 2, 3, device=cuda_device)], [8.0])
    with pytest.raises(RuntimeError):
        head_tail([torch.zeros(1, 64, 2, 2)], [torch.zeros(1, 3, 2, 2)], [8.0])
    with pytest.raises(TypeError):
        head_tail([t.double() for t in b], [torch.zeros(1, 3, 2, 2, device=cuda_device).double()], [8.0])
    # fp16 (the reference's default autocast dtype, train_model.py:240-246) is a plain 2-byte copy, forward and backward
    g = torch.Generator().manual_seed(3)
    bx = torch.randn(2, 64, 3, 5, generator=g).half().to(cuda_device).requires_grad_(True)
    cx = torch.randn(2, 3, 3, 5, generator=g).half().to(cuda_device).requires_grad_(True)
    x, anc, st = head_tail([bx], [cx], [8.0])
    assert x.dtype == torch.float16 and torch.equal(x, torch.cat((bx, cx), 1).view(2, 67, -1))
    w = torch.randn(x.shape, generator=g).half().to(cuda_device)
    (x * w).sum().backward()
    assert torch.equal(bx.grad.view(2, 64, -1), w[:, :64]) and torch.equal(cx.grad.view(2, 3, -1), w[:, 64:])
