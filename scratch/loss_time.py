"""Timing driver (not part of the product): ms/step of the nearest-centre step on the bench inputs.
   python scratch/loss_time.py [steps]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from custom_yolo_implmentation_b200.model.losses import fused_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
def run(tag, n, imgsz, gmax, seed, dt):
    preds, gts, anchors, strides = syn.make_loss_inputs(n, 80, imgsz, gmax, seed, dtype=dt)
    preds = preds.to(dev); anchors = anchors.float().to(dev); strides = strides.float().to(dev)
    gt, off, counts = pack_gt([g.to(dev) for g in gts], dev)
    f = lambda: fused_loss(preds, gt, off, max(counts), anchors, strides, 80, 1.0, 1.5)
    for _ in range(5): out, grad, _ = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): out, grad, _ = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    by = 2 * preds.numel() * preds.element_size()
    print(f"{tag:10s} {ms*1e3:8.1f} us/step  {by/ms/1e6:7.0f} GB/s  frac {by/ms/1e6/6534.8:.3f}  loss {out[0].item():.6f} stats {out.tolist()[3:]}", flush=True)
run('cfg2_f32', 128, 640, 100, 1234, torch.float32)
run('cfg2_bf16', 128, 640, 100, 1234, torch.bfloat16)
run('cfg5_bf16', 32, 1280, 300, 1240, torch.bfloat16)
