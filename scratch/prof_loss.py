"""Profiling driver (not part of the product): a few steps of one workload.
   python scratch/prof_loss.py loss|loss_bf16_cfg5|tal|nms"""
import sys, torch
sys.path.insert(0, '/root/repo')
from custom_yolo_implmentation_b200.model.losses import fused_loss, fused_tal_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.utils.model_utils import batched_nms_raw
dev = torch.device('cuda:0')
what = sys.argv[1] if len(sys.argv) > 1 else 'loss'
dt = torch.bfloat16 if 'bf16' in what else torch.float32
if what.startswith('loss') or what.startswith('tal'):
    n, nc = (32, 80) if 'cfg5' in what else (128, 80)
    imgsz, gmax = (1280, 300) if 'cfg5' in what else (640, 100)
    preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, 1236, dtype=dt)
    preds = preds.to(dev); anchors = anchors.float().to(dev); strides = strides.float().to(dev)
    gt, off, counts = pack_gt([g.to(dev) for g in gts], dev)
    for _ in range(4):
        if what.startswith('tal'):
            out, grad, _ = fused_tal_loss(preds, gt, off, anchors, strides, nc, 1.5, 1.0, 1.5)
        else:
            out, grad, _ = fused_loss(preds, gt, off, max(counts), anchors, strides, nc, 1.0, 1.5)
    torch.cuda.synchronize(); print('loss', out[0].item())
else:
    y = syn.make_nms_input(64, 80, 640, 2024).to(dev)
    for _ in range(3):
        rows, cnt, _ = batched_nms_raw(y, 0.001, 0.7, 300, 80)
    torch.cuda.synchronize(); print('kept', int(cnt.min()))
