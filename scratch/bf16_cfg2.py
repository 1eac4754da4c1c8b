"""cfg2 shape with bf16 head outputs: fused loss step time."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200.model.losses import fused_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
for dt in (torch.float32, torch.bfloat16):
    preds, gts, anchors, strides = syn.make_loss_inputs(128, 80, 640, 100, 1236, dtype=dt)
    x = preds.to(dev); a = anchors.float().to(dev); s = strides.float().to(dev)
    gt, off, counts = pack_gt([g.to(dev) for g in gts], dev)
    f = lambda: fused_loss(x, gt, off, max(counts), a, s, 80, 1.0, 1.5)
    for _ in range(5): out, grad, _ = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): out, grad, _ = f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 100
    nbytes = 2 * x.numel() * x.element_size()
    print(f'{dt}: {ms*1e3:.1f} us/step  {128/ms*1e3:.0f} img/s  {nbytes/ms/1e6:.0f} GB/s ({nbytes/ms/1e6/6534.8*100:.0f}% of 6534.8)  loss {out[0].item():.6f}')
