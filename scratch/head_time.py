"""Time the one-launch head tail against torch's cats (forward + backward), N=128, 640 px, nc=80."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from custom_yolo_implmentation_b200.model.head import gather_levels

dev = torch.device("cuda:0")
for dtype in (torch.float32, torch.bfloat16):
    shapes = [(80, 80), (40, 40), (20, 20)]
    N = 128
    box = [torch.randn(N, 64, h, w, device=dev, dtype=dtype).requires_grad_(True) for h, w in shapes]
    cls = [torch.randn(N, 80, h, w, device=dev, dtype=dtype).requires_grad_(True) for h, w in shapes]
    g = torch.randn(N, 144, 8400, device=dev, dtype=dtype)

    def ours():
        x = gather_levels(box, cls)
        x.backward(g)
        for t in box + cls: t.grad = None

    def ref():
        lv = [torch.cat((b, c), 1) for b, c in zip(box, cls)]
        x = torch.cat([l.view(N, 144, -1) for l in lv], 2)
        x.backward(g)
        for t in box + cls: t.grad = None

    def ours_fwd():
        with torch.no_grad(): gather_levels(box, cls)

    def ref_fwd():
        with torch.no_grad():
            lv = [torch.cat((b, c), 1) for b, c in zip(box, cls)]
            torch.cat([l.view(N, 144, -1) for l in lv], 2)

    for name, f in (("ours fwd", ours_fwd), ("torch fwd", ref_fwd), ("ours fwd+bwd", ours), ("torch fwd+bwd", ref)):
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        nbytes = g.numel() * g.element_size()
        print(f"{dtype} {name:14s} {ms*1e3:8.1f} us   ({nbytes/1e6:.0f} MB tensor, {2*nbytes/ms/1e6:.0f} GB/s per pass-equivalent)")
