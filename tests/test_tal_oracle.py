"""Known-answer tests of the task-aligned oracle (oracle/tal_oracle.py).  Nothing upstream pins this
tier (the reference has no TAL/CIoU/BCE), so the oracle is pinned by cases small enough to compute by
hand with scalar Python arithmetic."""
import math

import torch

from oracle import tal_oracle as T


def _preds_from_ltrb(ltrb, cls_logits):
    """(A,4) integer ltrb bins + (A,nc) logits -> (1, 64+nc, A) head output whose DFL expectation is ~exactly ltrb."""
    a, nc = cls_logits.shape
    x = torch.full((1, 64 + nc, a), -40.0)
    for i in range(a):
        for k in range(4):
            x[0, 16 * k + int(ltrb[i][k]), i] = 40.0
    x[0, 64:, :] = cls_logits.t()
    return x


def _ciou_scalar(b1, b2, eps=1e-7):
    x1, y1, x2, y2 = b1; u1, v1, u2, v2 = b2
    w1, h1, w2, h2 = x2 - x1, y2 - y1 + eps, u2 - u1, v2 - v1 + eps
    inter = max(min(x2, u2) - max(x1, u1), 0) * max(min(y2, v2) - max(y1, v1), 0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw, ch = max(x2, u2) - min(x1, u1), max(y2, v2) - min(y1, v1)
    c2 = cw * cw + ch * ch + eps
    rho2 = ((u1 + u2 - x1 - x2) ** 2 + (v1 + v2 - y1 - y2) ** 2) / 4
    v = 4 / math.pi ** 2 * (math.atan(w2 / h2) - math.atan(w1 / h1)) ** 2
    alpha = v / (v - iou + 1 + eps)
    return iou - (rho2 / c2 + v * alpha)


def test_ciou_identical_disjoint_and_scalar_agreement():
    b = torch.tensor([[10.0, 20.0, 50.0, 80.0]])
    assert abs(T.ciou(b, b).item() - 1.0) < 1e-6
    far = torch.tensor([[200.0, 300.0, 240.0, 360.0]])
    assert T.ciou(b, far).item() < 0                          # no overlap: only the distance penalty
    g = torch.Generator().manual_seed(3)
    p = torch.rand(50, 2, generator=g) * 100
    b1 = torch.cat((p, p + 5 + torch.rand(50, 2, generator=g) * 60), 1)
    q = torch.rand(50, 2, generator=g) * 100
    b2 = torch.cat((q, q + 5 + torch.rand(50, 2, generator=g) * 60), 1)
    got = T.ciou(b1, b2)
    for i in range(50):
        assert abs(got[i].item() - _ciou_scalar(b1[i].tolist(), b2[i].tolist())) < 2e-6


def test_known_answer_two_gt_six_anchors():
    """One 3x2 grid of stride-8 anchors (centres 4,12,20 x 4,12), two GTs, top-k = 2."""
    anchors = torch.tensor([[0.5, 1.5, 2.5, 0.5, 1.5, 2.5], [0.5, 0.5, 0.5, 1.5, 1.5, 1.5]])
    strides = torch.full((1, 6), 8.0)
    # every anchor predicts the box  centre +- (1,1) cells = 16x16 px around its centre
    ltrb = [[1, 1, 1, 1]] * 6
    logit = lambda p: math.log(p / (1 - p))
    cls = torch.full((6, 2), logit(0.1))
    cls[0, 0] = logit(0.9); cls[1, 0] = logit(0.5); cls[3, 0] = logit(0.8)       # class 0 scores
    cls[2, 1] = logit(0.6); cls[5, 1] = logit(0.7); cls[1, 1] = logit(0.95)      # class 1 scores
    preds = _preds_from_ltrb(ltrb, cls)
    # GT0 class 0 covers anchors 0,1,3,4 (x in (0,16), y in (0,16)); GT1 class 1 covers anchors 1,2,4,5 (x in (8,24))
    gts = [torch.tensor([[8.0, 8.0, 16.0, 16.0, 0.0], [16.0, 8.0, 16.0, 16.0, 1.0]])]
    tr = T.tal_forward(preds, gts, anchors, strides, 2, topk=2, lambda_box=1.0, lambda_cls=1.0, lambda_dfl=1.0)
    box = lambda cx, cy: (cx - 8, cy - 8, cx + 8, cy + 8)
    ctr = [(4, 4), (12, 4), (20, 4), (4, 12), (12, 12), (20, 12)]
    g0, g1 = (0, 0, 16, 16), (8, 0, 24, 16)
    ov = lambda g, i: max(_ciou_scalar(g, box(*ctr[i])), 0.0)
    m0 = {i: math.sqrt(s) * ov(g0, i) ** 6 for i, s in ((0, 0.9), (1, 0.5), (3, 0.8), (4, 0.1))}
    m1 = {i: math.sqrt(s) * ov(g1, i) ** 6 for i, s in ((1, 0.95), (2, 0.6), (4, 0.1), (5, 0.7))}
    top0 = sorted(m0, key=lambda i: (-m0[i], i))[:2]
    top1 = sorted(m1, key=lambda i: (-m1[i], i))[:2]
    assert sorted(tr.topk_anchor[0][0].tolist()) == sorted(top0) == [0, 3]
    assert sorted(tr.topk_anchor[0][1].tolist()) == sorted(top1) == [1, 5]
    assert tr.assigned_gt[0].tolist() == [0, 1, -1, 0, -1, 1] and tr.num_fg == 4
    # normalised targets: metric * max_overlap / max_metric per GT
    t = [0.0] * 6
    for g, m, top in ((g0, m0, top0), (g1, m1, top1)):
        mx_m, mx_o = max(m[i] for i in top), max(ov(g, i) for i in top)
        for i in top:
            t[i] = m[i] * mx_o / (mx_m + 1e-9)
    assert torch.allclose(tr.target_score[0], torch.tensor(t), rtol=1e-5, atol=1e-7)
    tss = max(sum(t), 1.0)
    # class loss: BCE over all 12 cells
    bce = 0.0
    for i in range(6):
        for c in range(2):
            x = cls[i, c].item()
            tgt = t[i] if (tr.assigned_gt[0, i].item() == c) else 0.0
            bce += max(x, 0) - x * tgt + math.log1p(math.exp(-abs(x)))
    assert abs(tr.cls.item() - bce / tss) < 1e-5 * bce / tss
    # varifocal weighting of the same cells: alpha * p^gamma on background, the target score on the positive cell
    trv = T.tal_forward(preds, gts, anchors, strides, 2, topk=2, lambda_box=1.0, lambda_cls=1.0, lambda_dfl=1.0,
                        cls_loss="vfl", vfl_alpha=0.75, vfl_gamma=2.0)
    vfl = 0.0
    for i in range(6):
        for c in range(2):
            x = cls[i, c].item()
            pos = tr.assigned_gt[0, i].item() == c
            tgt = t[i] if pos else 0.0
            cell = max(x, 0) - x * tgt + math.log1p(math.exp(-abs(x)))
            p = 1.0 / (1.0 + math.exp(-x))
            vfl += cell * (tgt if pos else 0.75 * p ** 2)
    assert abs(trv.cls.item() - vfl / tss) < 1e-5 * vfl / tss
    assert trv.assigned_gt.equal(tr.assigned_gt) and abs(trv.box.item() - tr.box.item()) < 1e-7
    # box loss
    bl = sum((1 - _ciou_scalar(box(*ctr[i]), (g0 if tr.assigned_gt[0, i] == 0 else g1))) * t[i] for i in range(6) if t[i] > 0)
    assert abs(tr.box.item() - bl / tss) < 1e-5
    assert tr.dfl.item() > 0 and abs(tr.total.item() - (tr.box + tr.cls + tr.dfl).item()) < 1e-6


def test_conflict_goes_to_the_larger_overlap_and_empty_images_are_fine():
    anchors = torch.tensor([[0.5, 1.5], [0.5, 0.5]])
    strides = torch.full((1, 2), 8.0)
    cls = torch.zeros(2, 1)
    preds = _preds_from_ltrb([[1, 1, 1, 1], [1, 1, 1, 1]], cls)
    # both GTs contain anchor 0 only (centre (4,4)); GT1 matches its predicted box (-4,-4,12,12) better
    gts = [torch.tensor([[4.0, 4.0, 7.0, 7.0, 0.0], [4.0, 4.0, 15.0, 15.0, 0.0]]), torch.zeros(0, 5)]
    preds = torch.cat((preds, preds), 0)
    tr = T.tal_forward_backward(preds, gts, anchors, strides, 1, topk=3)
    assert tr.assigned_gt[0].tolist() == [1, -1] and tr.assigned_gt[1].tolist() == [-1, -1]
    assert tr.grad.shape == preds.shape and torch.isfinite(tr.grad).all()
    assert tr.grad[1, :64].abs().sum().item() == 0            # no foreground in the empty image: no box gradient
    assert tr.grad[1, 64:].abs().sum().item() > 0             # but its class logits still see the BCE
