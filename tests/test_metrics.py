"""Validation bookkeeping: the oracle against the reference's golden counters (CPU), the CUDA kernel
against both (GPU).  All quantities are integers: bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.metrics_oracle import MetricsOracle


def _cases(z):
    n = int(z["meta"][0])
    for i in range(n):
        p, m = int(z["pred_count"][i]), int(z["target_count"][i])
        yield (torch.from_numpy(z["preds"][i, :p].copy()), torch.from_numpy(z["targets"][i, :m].copy()),
               torch.from_numpy(z["scores"][i, :p].copy()) if bool(z["use_scores"][i]) else None)


def _expected(z):
    nc = int(z["meta"][1])
    out = np.zeros(8 + 4 * nc, np.int64)
    out[:5] = z["totals"]
    out[8:] = np.concatenate([z["class_tp"], z["class_fp"], z["class_fn"], z["class_gt"]]).astype(np.int64)
    return out


@pytest.mark.parametrize("name", ["metrics_a", "metrics_b"])
def test_metrics_oracle_matches_reference(name):
    z = load_golden(name)
    o = MetricsOracle(int(z["meta"][1]), float(z["thr"]))
    for pr, t, sc in _cases(z):
        o.update(pr, t, sc, 0.3)
    assert np.array_equal(o.vector(), _expected(z))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["metrics_a", "metrics_b"])
def test_metrics_cuda_matches_reference(name, cuda_device):
    from custom_yolo_implmentation_b200.model.losses import pack_gt
    from custom_yolo_implmentation_b200.training.metrics import DetectionMetrics
    z = load_golden(name)
    nc = int(z["meta"][1])
    dev = cuda_device
    # (1) the reference's calling pattern: one update() per image
    mt = DetectionMetrics(nc, float(z["thr"]))
    for pr, t, sc in _cases(z):
        mt.update(pr.to(dev), t.to(dev), pred_scores=None if sc is None else sc.to(dev), score_threshold=0.3)
    assert np.array_equal(mt._host().numpy(), _expected(z))
    res = mt.compute()
    assert np.allclose([res["precision"], res["recall"], res["f1_score"], res["mAP"]], z["compute"], rtol=1e-6, atol=1e-9)
    assert (mt.true_positives, mt.false_positives, mt.false_negatives) == tuple(int(v) for v in z["totals"][:3])
    assert torch.equal(mt.class_tp, torch.from_numpy(z["class_tp"])) and mt.get_class_metrics(0)["ground_truths"] == int(z["class_gt"][0])
    # (2) the whole batch in one launch (scores of the unfiltered images set to 1 so that they pass)
    mb = DetectionMetrics(nc, float(z["thr"]))
    tg = [t for _, t, _ in _cases(z)]
    gt, off, counts = pack_gt([t.to(dev) for t in tg], dev)
    rows = torch.from_numpy(z["preds"]).to(dev)
    sc = torch.from_numpy(z["scores"]).to(dev).clone()
    sc[~torch.from_numpy(z["use_scores"]).to(dev)] = 1.0
    mb.update_batch(rows, torch.from_numpy(z["pred_count"]).to(dev), gt, off, max(counts), sc, 0.3,
                    skip_empty_targets=False)               # the golden pins update() on EVERY image
    assert np.array_equal(mb._host().numpy(), _expected(z))
    mb.reset()
    assert mb.true_positives == 0 and mb.compute()["mAP"] == 0.0


@pytest.mark.gpu
def test_metrics_cuda_matches_oracle_at_validation_size(cuda_device):
    """64 images x 100 predictions x up to 100 targets, fed straight from decode_predictions_raw."""
    from custom_yolo_implmentation_b200.model.losses import pack_gt
    from custom_yolo_implmentation_b200.training.metrics import DetectionMetrics, compute_average_iou
    from custom_yolo_implmentation_b200.training.train_model import decode_predictions_raw
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    dev = cuda_device
    n, nc = 64, 80
    anchors, strides = syn.anchor_grid(640)
    preds = syn.make_preds(n, nc, anchors.shape[1], 77, cls_mean=-1.0, cls_std=1.5)
    gts = syn.make_gt(n, nc, 640, 100, 78)
    rows, count, _ = decode_predictions_raw(preds.to(dev), anchors.to(dev), strides.to(dev), 0.25, 100, nc)
    # plant some true positives: copy a few GT boxes (right class) into the prediction rows
    for b in range(0, n, 3):
        k = min(5, gts[b].shape[0], int(count[b]))
        if k:
            rows[b, :k] = gts[b][:k].to(dev)
        if gts[b].shape[0] > 70 and int(count[b]) > 12:       # ... and some of the targets beyond the 64 a warp keeps in registers
            rows[b, 8:11] = gts[b][-3:].to(dev)
            rows[b, 11] = gts[b][-1].to(dev)                  # a duplicate: its target is already consumed -> false positive
    gt, off, counts = pack_gt([g.to(dev) for g in gts], dev)
    mt = DetectionMetrics(nc, 0.5)
    mt.update_batch(rows, count, gt, off, max(counts), skip_empty_targets=False)
    o = MetricsOracle(nc, 0.5)
    rows_h, count_h = rows.cpu(), count.cpu()
    for b in range(n):
        o.update(rows_h[b, : int(count_h[b])], gts[b])
    assert np.array_equal(mt._host().numpy(), o.vector())
    assert mt.true_positives > 0
    # the reference's validation loop (train_model.py:326-328) skips images without targets: that is the default,
    # also with a PackedGT straight from the collate function (host memory: moved, not dereferenced)
    assert any(g.shape[0] == 0 for g in gts)
    from custom_yolo_implmentation_b200.model.losses import pack_gt_host
    ml = DetectionMetrics(nc, 0.5)
    ml.update_batch(rows, count, pack_gt_host(gts))
    ol = MetricsOracle(nc, 0.5)
    for b in range(n):
        if gts[b].numel() > 0:
            ol.update(rows_h[b, : int(count_h[b])], gts[b])
    assert np.array_equal(ml._host().numpy(), ol.vector())
    assert ml.false_positives < mt.false_positives
    avg = compute_average_iou([rows[b, : int(count_h[b]), :4] for b in range(4)], [g[:, :4].to(dev) for g in gts[:4]])
    assert 0.0 <= avg <= 1.0
