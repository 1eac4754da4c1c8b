"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run only in the build container, where the reference is mounted read-only:

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (pure Python, /root/reference) cannot travel to the GPU box, so its outputs on a
handful of small seeded inputs are committed here together with the inputs themselves.  The
fixtures pin (a) the CPU oracle under ``oracle/`` and (b) the CUDA path, both in ``tests/``.

What is called, unmodified, from /root/reference:
  src.model.losses.YoloDFLQFLoss / bbox_iou / quality_focal_loss / distribution_focal_loss
  src.utils.model_utils.non_max_suppression / make_anchors / dist2bbox / box_iou / xywh2xyxy
  src.model.model_blocks.DFL
  src.training.train_model.decode_predictions
  src.training.metrics.box_iou_batch
  src.model.head.Head  (forward; the six conv-tower outputs it concatenates are captured with hooks)
The only intervention is freezing ``src.utils.model_utils.time`` so the reference's wall-clock
NMS abort (model_utils.py:212, :275-277) cannot drop images on a slow machine.  The matched
anchor indices, which the reference computes but does not return, are recovered by re-running
its own two lines (losses.py:214-215) on its own decoded centres (losses.py:155-188).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


syn = _load("_synthetic", os.path.join(ROOT, "custom-yolo-implmentation_b200", "utils", "synthetic.py"))

sys.path.insert(0, REF)
import src.model.losses as ref_losses            # noqa: E402
import src.utils.model_utils as ref_utils        # noqa: E402
import src.model.model_blocks as ref_blocks      # noqa: E402
import src.training.train_model as ref_train     # noqa: E402
import src.training.metrics as ref_metrics       # noqa: E402
import src.model.head as ref_head                # noqa: E402

ref_utils.time = types.SimpleNamespace(time=lambda: 0.0)     # freeze the NMS abort clock (Q8)


def pack_gt(gts):
    gmax = max(1, max(g.shape[0] for g in gts))
    out = np.zeros((len(gts), gmax, 5), np.float32)
    cnt = np.zeros((len(gts),), np.int32)
    for i, g in enumerate(gts):
        out[i, : g.shape[0]] = g.numpy()
        cnt[i] = g.shape[0]
    return out, cnt


def pack_ragged(tensors, width, dtype):
    k = max(1, max(t.shape[0] for t in tensors))
    out = np.zeros((len(tensors), k) + ((width,) if width else ()), dtype)
    cnt = np.zeros((len(tensors),), np.int32)
    for i, t in enumerate(tensors):
        out[i, : t.shape[0]] = t.detach().numpy()
        cnt[i] = t.shape[0]
    return out, cnt


def ref_matched_idx(preds, gts, anchors, strides, reg_max=16):
    """losses.py:142-188 + :211-215 re-executed verbatim on the reference's own tensors."""
    p = preds.float().transpose(1, 2)
    anc = anchors.transpose(0, 1)
    st = strides.transpose(0, 1)
    B, A, _ = p.shape
    pd = p[:, :, : 4 * reg_max].view(B, A, 4, reg_max).softmax(3)
    ltrb = torch.sum(pd * torch.arange(reg_max, dtype=p.dtype), dim=3)
    x1 = (anc[None, :, 0] - ltrb[:, :, 0]) * st[None, :, 0]
    y1 = (anc[None, :, 1] - ltrb[:, :, 1]) * st[None, :, 0]
    x2 = (anc[None, :, 0] + ltrb[:, :, 2]) * st[None, :, 0]
    y2 = (anc[None, :, 1] + ltrb[:, :, 3]) * st[None, :, 0]
    xywh = torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], dim=2)
    out, ious = [], []
    for b in range(B):
        g = gts[b]
        if g.numel() == 0:
            out.append(torch.zeros(0, dtype=torch.long)); ious.append(torch.zeros(0)); continue
        idx = torch.cdist(g[:, 0:2].to(p.dtype), xywh[b, :, 0:2]).argmin(dim=1)
        out.append(idx)
        ious.append(ref_losses.bbox_iou(xywh[b][idx], g[:, 0:4].to(p.dtype)))
    return out, ious, xywh


ONLY = set(sys.argv[1:])          # optional: regenerate just the named fixtures


def _skip(name):
    return bool(ONLY) and name not in ONLY


def loss_case(name, n, nc, imgsz, gmax, seed, dtype=torch.float32, conflict=0.0, store_inputs=True, sample=None):
    if _skip(name):
        return
    preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, seed, dtype=dtype, conflict_frac=conflict)
    leaf = preds.clone().requires_grad_(True)
    crit = ref_losses.YoloDFLQFLoss(num_classes=nc)
    loss, parts = crit(leaf, gts, anchors, strides)
    loss.backward()
    idx, ious, xywh = ref_matched_idx(preds, gts, anchors, strides)
    gt_pad, gt_cnt = pack_gt(gts)
    idx_pad, _ = pack_ragged(idx, 0, np.int64)
    iou_pad, _ = pack_ragged(ious, 0, np.float32)
    grad = leaf.grad
    d = dict(meta=np.array([n, nc, imgsz, gmax, seed, int(dtype == torch.bfloat16), int(round(conflict * 100))], np.int64),
             total_loss=np.float32(parts["total_loss"]), box_loss=np.float32(parts["box_loss"]),
             cls_loss=np.float32(parts["cls_loss"]), loss_tensor=loss.detach().numpy(),
             gt=gt_pad, gt_count=gt_cnt, idx=idx_pad, iou=iou_pad)
    if store_inputs:
        as_np = (lambda t: t.view(torch.int16).numpy()) if dtype == torch.bfloat16 else (lambda t: t.numpy())
        d.update(preds=as_np(preds), anchors=as_np(anchors), strides=as_np(strides), grad=as_np(grad.detach()),
                 pred_xywh=xywh.numpy())
    else:
        g = grad.detach().float().flatten()
        pos = torch.arange(0, g.numel(), sample)
        d.update(grad_sample_stride=np.int64(sample), grad_sample=g[pos].numpy(),
                 grad_abs_sum=np.float64(g.double().abs().sum().item()), grad_sum=np.float64(g.double().sum().item()),
                 grad_absmax=np.float32(g.abs().max().item()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    dup = sum(len(i) - len(i.unique()) for i in idx)
    print(f"{name}: loss {parts['total_loss']:.6f} box {parts['box_loss']:.6f} cls {parts['cls_loss']:.6f} "
          f"GT {int(gt_cnt.sum())} duplicate-anchor GTs {dup}")


def rank_losses_case(name, world, n, nc, imgsz, gmax, seed0):
    """cfg3 (8 ranks x cfg2): the reference's loss scalars and matched anchors of every rank's shard of bench.py
    (rank r draws its batch with seed seed0 + r), so that the multi-GPU run can be checked rank by rank on the box."""
    if _skip(name):
        return
    tot, box, cls, idx_all, cnt_all = [], [], [], [], []
    for r in range(world):
        preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, seed0 + r)
        with torch.no_grad():
            loss, parts = ref_losses.YoloDFLQFLoss(num_classes=nc)(preds, gts, anchors, strides)
        idx, _, _ = ref_matched_idx(preds, gts, anchors, strides)
        tot.append(parts["total_loss"]); box.append(parts["box_loss"]); cls.append(parts["cls_loss"])
        idx_all.append(torch.cat(idx).numpy().astype(np.int32)); cnt_all.append(np.array([g.shape[0] for g in gts], np.int32))
        print(f"{name} rank {r}: loss {parts['total_loss']:.6f} GT {int(cnt_all[-1].sum())}")
    k = max(len(i) for i in idx_all)
    idx_pad = np.full((world, k), -1, np.int32)
    for r, i in enumerate(idx_all):
        idx_pad[r, : len(i)] = i
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array([world, n, nc, imgsz, gmax, seed0], np.int64),
                        total_loss=np.array(tot, np.float32), box_loss=np.array(box, np.float32),
                        cls_loss=np.array(cls, np.float32), idx=idx_pad, gt_count=np.stack(cnt_all))


def nms_case(name, n, nc, imgsz, seed, conf, iou, max_det=300, agnostic=False, classes=None, dense=False, nm=0,
             multi_label=False, allow_ties=False, store_inputs=True):
    if _skip(name):
        return
    x = syn.make_nms_input(n, nc, imgsz, seed, dense_uniform=dense)
    if nm:
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.cat((x, torch.randn(n, nm, x.shape[2], generator=g)), 1)
    for b in range(n):      # bit-exact keep lists need unique scores per image (Q10)
        best = x[b, 4:4 + nc].flatten() if multi_label else x[b, 4:4 + nc].amax(0)
        assert allow_ties or best.unique().numel() == best.numel(), "duplicate best scores; pick another seed"
    out = ref_utils.non_max_suppression(x, conf_thres=conf, iou_thres=iou, classes=classes, agnostic=agnostic,
                                        multi_label=multi_label, max_det=max_det, nc=nc)
    rows, cnt = pack_ragged(out, 6 + nm, np.float32)
    # store_inputs=False: the test regenerates the prediction from (n, nc, imgsz, seed) with utils/synthetic.py
    np.savez_compressed(os.path.join(HERE, name + ".npz"), prediction=x.numpy() if store_inputs else np.zeros(0, np.float32),
                        rows=rows, count=cnt,
                        meta=np.array([n, nc, imgsz, seed, max_det, int(agnostic), nm], np.int64),
                        conf=np.float64(conf), iou=np.float64(iou),
                        classes=np.array(classes if classes is not None else [], np.int64))
    print(f"{name}: kept per image {cnt.tolist()}")


def decode_case(name, n, nc, imgsz, seed, conf, top_k, cls_mean):
    if _skip(name):
        return
    anchors, strides = syn.anchor_grid(imgsz)
    preds = syn.make_preds(n, nc, anchors.shape[1], seed, cls_mean=cls_mean, cls_std=1.5)
    out = ref_train.decode_predictions(preds, anchors, strides, conf_threshold=conf, top_k=top_k, num_classes=nc)
    rows, cnt = pack_ragged(out, 5, np.float32)
    # inference-side decode (model_builder.py:123-133) on the same tensor
    dfl = ref_blocks.DFL(16)
    with torch.no_grad():
        ltrb = dfl(preds[:, :64, :])
        box = ref_utils.dist2bbox(ltrb, anchors.unsqueeze(0), xywh=True, dim=1) * strides
        box_xyxy = ref_utils.dist2bbox(ltrb, anchors.unsqueeze(0), xywh=False, dim=1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), preds=preds.numpy(), anchors=anchors.numpy(),
                        strides=strides.numpy(), rows=rows, count=cnt, dfl_ltrb=ltrb.numpy(), box_xywh=box.numpy(),
                        box_xyxy_grid=box_xyxy.numpy(), meta=np.array([n, nc, imgsz, seed, top_k], np.int64),
                        conf=np.float64(conf))
    print(f"{name}: rows per image {cnt.tolist()}")


def inference_case(name, n, nc, imgsz, seed, conf, iou, cls_mean):
    """Model.inference after the network, the reference's own lines (src/model/model_builder.py:123-139) on a seeded
    head output: split, DFL, dist2bbox(xywh), * strides, cat, non_max_suppression -- RAW logits as scores."""
    if _skip(name):
        return
    anchors, strides = syn.anchor_grid(imgsz)
    x = syn.make_preds(n, nc, anchors.shape[1], seed, cls_mean=cls_mean, cls_std=1.5)
    dfl = ref_blocks.DFL(16)
    with torch.no_grad():
        box, cls = x.split((64, nc), 1)                                  # :123
        box = dfl(box)                                                   # :127
        box = ref_utils.dist2bbox(box, anchors.unsqueeze(0), xywh=True, dim=1)   # :130
        box = box * strides                                              # :133
        y = torch.cat((box, cls), 1)                                     # :136
        for b in range(n):
            best = y[b, 4:].amax(0)
            assert best.unique().numel() == best.numel(), "duplicate best scores; pick another seed"
        out = ref_utils.non_max_suppression(y, conf_thres=conf, iou_thres=iou, nc=nc)   # :139
    rows, cnt = pack_ragged(out, 6, np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x.numpy(), anchors=anchors.numpy(), strides=strides.numpy(),
                        rows=rows, count=cnt, meta=np.array([n, nc, imgsz, seed], np.int64), conf=np.float64(conf),
                        iou=np.float64(iou))
    print(f"{name}: kept per image {cnt.tolist()}")


def metrics_case(name, seed, n_images, nc, thr):
    if _skip(name):
        return
    """DetectionMetrics of the reference on seeded predictions/targets (incl. empty images, score filter)."""
    g = torch.Generator().manual_seed(seed)
    mt = ref_metrics.DetectionMetrics(nc, iou_threshold=thr)
    preds, tgts, scores = [], [], []
    for i in range(n_images):
        m = int(torch.randint(0, 9, (1,), generator=g))
        p = int(torch.randint(0, 14, (1,), generator=g))
        if i == 1: m = 0
        if i == 2: p = 0
        t = torch.cat((torch.rand(m, 2, generator=g) * 200, 10 + torch.rand(m, 2, generator=g) * 60,
                       torch.randint(0, nc, (m, 1), generator=g).float()), 1)
        # predictions: jittered copies of targets (some right class, some wrong) + random boxes
        k = min(p, m)
        src = torch.randperm(m, generator=g)[:k] if m else torch.zeros(0, dtype=torch.long)
        near = t[src].clone()
        near[:, :4] += (torch.rand(k, 4, generator=g) - 0.5) * 12
        flip = torch.rand(k, generator=g) < 0.25
        near[flip, 4] = torch.randint(0, nc, (int(flip.sum()),), generator=g).float()
        rnd = torch.cat((torch.rand(p - k, 2, generator=g) * 200, 10 + torch.rand(p - k, 2, generator=g) * 60,
                         torch.randint(0, nc, (p - k, 1), generator=g).float()), 1)
        pr = torch.cat((near, rnd), 0)[torch.randperm(p, generator=g)] if p else torch.zeros(0, 5)
        sc = torch.rand(p, generator=g)
        mt.update(pr, t, pred_scores=sc if i % 2 == 0 else None, score_threshold=0.3)
        preds.append(pr); tgts.append(t); scores.append(sc if i % 2 == 0 else torch.ones(p))
    out = mt.compute()
    pp, pc = pack_ragged(preds, 5, np.float32)
    tt, tc = pack_ragged(tgts, 5, np.float32)
    ss, _ = pack_ragged(scores, 0, np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), preds=pp, pred_count=pc, targets=tt, target_count=tc, scores=ss,
                        use_scores=np.array([i % 2 == 0 for i in range(n_images)]), meta=np.array([n_images, nc], np.int64),
                        thr=np.float64(thr),
                        totals=np.array([mt.true_positives, mt.false_positives, mt.false_negatives, mt.total_predictions,
                                         mt.total_ground_truths], np.int64),
                        class_tp=mt.class_tp.numpy(), class_fp=mt.class_fp.numpy(), class_fn=mt.class_fn.numpy(),
                        class_gt=mt.class_gt_count.numpy(),
                        compute=np.array([out["precision"], out["recall"], out["f1_score"], out["mAP"]], np.float64))
    print(f"{name}: TP {mt.true_positives} FP {mt.false_positives} FN {mt.false_negatives}")


def helper_case(name, seed):
    if _skip(name):
        return
    g = torch.Generator().manual_seed(seed)
    m, c = 37, 11
    b1 = torch.cat((torch.rand(m, 2, generator=g) * 200, 5 + torch.rand(m, 2, generator=g) * 90), 1)
    b2 = b1 + (torch.rand(m, 4, generator=g) - 0.5) * 30
    b2[:, 2:].clamp_(min=1.0)
    q = torch.cat((torch.rand(23, 2, generator=g) * 200, 5 + torch.rand(23, 2, generator=g) * 90), 1)
    scores = torch.randn(m, c, generator=g) * 2 - 2
    target = torch.zeros(m, c)
    target[torch.arange(m), torch.randint(0, c, (m,), generator=g)] = torch.rand(m, generator=g)
    dist = torch.randn(m, 16, generator=g)
    tval = torch.rand(m, generator=g) * 14.99
    lv = [torch.zeros(1, 1, 6, 5), torch.zeros(1, 1, 3, 3), torch.zeros(1, 1, 2, 1)]
    anc, st = ref_utils.make_anchors(lv, [8, 16, 32], 0.5)
    c1, c2 = ref_utils.xywh2xyxy(b1), ref_utils.xywh2xyxy(q)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), b1=b1.numpy(), b2=b2.numpy(), q=q.numpy(), scores=scores.numpy(),
        target=target.numpy(), dist=dist.numpy(), tval=tval.numpy(),
        bbox_iou=ref_losses.bbox_iou(b1, b2).numpy(), qfl=ref_losses.quality_focal_loss(scores, target).numpy(),
        dfl=ref_losses.distribution_focal_loss(dist, tval).numpy(), xyxy=c1.numpy(),
        box_iou=ref_utils.box_iou(c1, c2).numpy(), box_iou_batch=ref_metrics.box_iou_batch(b1, q).numpy(),
        anchors=anc.numpy(), strides=st.numpy())
    print(f"{name}: helpers written")


def head_case(name, seed, n, nc, shapes, dtype=torch.float32):
    if _skip(name):
        return
    """Head.forward of the reference (head.py:77-121) on random feature maps; the outputs of its box / cls
    towers (what the tail concatenates) are captured with forward hooks."""
    torch.manual_seed(seed)
    filters = [8, 16, 32][: len(shapes)]
    head = ref_head.Head(nc=nc, filters=filters).eval()
    head.stride = torch.tensor([8.0, 16.0, 32.0][: len(shapes)])
    cap = {}
    for i in range(len(shapes)):
        head.box[i].register_forward_hook(lambda m, a, out, i=i: cap.__setitem__(("box", i), out.detach().clone()))
        head.cls[i].register_forward_hook(lambda m, a, out, i=i: cap.__setitem__(("cls", i), out.detach().clone()))
    feats = [torch.randn(n, f, h, w) for f, (h, w) in zip(filters, shapes)]
    with torch.no_grad():
        x, anchors, strides = head(list(feats))
    d = {"x": x.to(dtype).float().numpy(), "anchors": anchors.numpy(), "strides": strides.numpy(),
         "meta": np.array([n, nc, len(shapes)], np.int64), "shapes": np.array(shapes, np.int64)}
    for i in range(len(shapes)):
        d[f"box{i}"] = cap[("box", i)].to(dtype).float().numpy()
        d[f"cls{i}"] = cap[("cls", i)].to(dtype).float().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(f"{name}: x {tuple(x.shape)}")


if __name__ == "__main__":
    torch.set_num_threads(1)         # fixed reduction order for the stored reference sums
    loss_case("loss_small_fp32", 3, 6, 128, 10, 101)
    loss_case("loss_conflict_fp32", 3, 6, 128, 12, 202, conflict=0.4)
    loss_case("loss_small_bf16", 2, 6, 128, 10, 303, dtype=torch.bfloat16)
    loss_case("loss_nc171_fp32", 1, 171, 96, 5, 404)
    loss_case("loss_cfg1_summary", 16, 80, 640, 50, 1235, store_inputs=False, sample=4099)
    # the full cfg2 batch (N=128, <=100 GT/img, fp32) and four images of the cfg5 shape (1280x1280, <=300 GT/img, bf16
    # head outputs): summaries only, the tests regenerate the inputs from the seed
    loss_case("loss_cfg2_summary", 128, 80, 640, 100, 1236, store_inputs=False, sample=65537)
    loss_case("loss_cfg5_summary", 4, 80, 1280, 300, 1240, dtype=torch.bfloat16, store_inputs=False, sample=16411)
    rank_losses_case("loss_cfg3_ranks", 8, 128, 80, 640, 100, 1236)
    nms_case("nms_small", 3, 6, 160, 11, conf=0.001, iou=0.7)
    nms_case("nms_dense_maxdet", 2, 4, 160, 12, conf=0.25, iou=0.45, max_det=40, dense=True)
    nms_case("nms_agnostic", 2, 6, 160, 13, conf=0.05, iou=0.5, agnostic=True)
    nms_case("nms_classes", 2, 6, 160, 14, conf=0.3, iou=0.6, classes=[1, 4])
    nms_case("nms_few", 3, 6, 160, 15, conf=0.9, iou=0.3)
    nms_case("nms_multilabel", 3, 6, 160, 16, conf=0.05, iou=0.6, multi_label=True)
    nms_case("nms_multilabel_classes", 2, 6, 160, 19, conf=0.02, iou=0.5, multi_label=True, classes=[0, 2, 5], max_det=50)
    # 672 000 candidates > max_nms = 30 000; equal scores exist among them (the reference's argsort decides their order)
    nms_case("nms_multilabel_cap", 1, 80, 640, 18, conf=0.001, iou=0.7, multi_label=True, allow_ties=True,
             store_inputs=False)
    decode_case("decode_topk", 3, 6, 160, 21, conf=0.25, top_k=100, cls_mean=-1.0)
    decode_case("decode_sparse", 3, 6, 160, 22, conf=0.6, top_k=100, cls_mean=-4.0)
    inference_case("inference_post", 3, 6, 160, 23, conf=0.25, iou=0.45, cls_mean=-1.0)
    helper_case("helpers", 31)
    metrics_case("metrics_a", 41, 12, 5, 0.5)
    metrics_case("metrics_b", 42, 9, 3, 0.3)
    head_case("head_aligned", 51, 2, 5, [(8, 8), (4, 4), (2, 4)])
    head_case("head_ragged", 52, 3, 3, [(7, 5), (3, 3), (2, 1)])
