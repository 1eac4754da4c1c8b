"""Validation path timing: decode_predictions_raw (yb_val_decode) + DetectionMetrics.update_batch, N=64, 640 px, nc=80."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200.training.train_model import decode_predictions_raw
from custom_yolo_implmentation_b200.training.metrics import DetectionMetrics
dev = torch.device('cuda:0')
for dt in (torch.float32, torch.bfloat16):
    preds, gts, anchors, strides = syn.make_loss_inputs(64, 80, 640, 50, 7, dtype=dt)
    x = preds.to(dev); a = anchors.to(dev); s = strides.to(dev)
    gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev)
    m = DetectionMetrics(num_classes=80)
    def dec(): return decode_predictions_raw(x, a, s, 0.25, 100, 80)
    rows, cnt, _ = dec()
    def met(): m.update_batch(rows, cnt, gt, off, max(counts))
    for name, f in (("val decode", dec), ("metrics update_batch", met)):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        nbytes = x.numel() * x.element_size()
        print(f'{dt} {name:22s} {ms*1e3:8.1f} us/call   ({nbytes/1e6:.0f} MB input -> {nbytes/ms/1e6:.0f} GB/s)   kept {cnt[:3].tolist()}')
