"""Does a programmatic dependent launch of fused_main_kernel cost the host-synchronised API path anything?"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from custom_yolo_implmentation_b200.model.losses import YoloDFLQFLoss, fused_loss, pack_gt_host
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.utils.model_utils import make_anchors
from custom_yolo_implmentation_b200 import _cabi
import custom_yolo_implmentation_b200.model.losses as L
dev = torch.device("cuda:0")
n, nc = 128, 80
preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, 640, 100, 1)
preds = preds.to(dev); gts_d = [g.to(dev) for g in gts]; anchors = anchors.to(dev); strides = strides.to(dev)
packed = pack_gt_host(gts, pin_memory=False).to(dev)
crit = YoloDFLQFLoss(num_classes=nc)
x = preds.clone().requires_grad_(True)
orig = L.fused_loss
def run(extra, gt_arg, read, k=200):
    def fl(*a, **kw):
        kw["flags"] = kw.get("flags", 0) | extra
        return orig(*a, **kw)
    L.fused_loss = fl
    def step():
        x.grad = None
        loss, parts = crit(x, gt_arg, anchors, strides)
        loss.backward()
        return parts["total_loss"] if read else None
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(k): step()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k, th / k * 1e3
for name, extra in (("pdl", 0), ("plain", _cabi.YB_LOSS_NO_PDL), ("pdl", 0), ("plain", _cabi.YB_LOSS_NO_PDL)):
    for gname, g in (("packed", packed), ("list", gts_d)):
        for read in (True, False):
            print(name, gname, "read" if read else "noread", "gpu ms/step %.4f host ms/step %.4f" % run(extra, g, read), flush=True)
