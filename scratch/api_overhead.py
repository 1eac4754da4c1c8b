"""Where the time of one YoloDFLQFLoss.forward + backward goes on the host (not product code)."""
import sys, time, torch, cProfile, pstats
sys.path.insert(0, '/root/repo')
from custom_yolo_implmentation_b200.model.losses import YoloDFLQFLoss, pack_gt_host
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
n, nc = 128, 80
preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, 640, 100, 1236)
x = preds.to(dev).requires_grad_(True); anchors = anchors.to(dev); strides = strides.to(dev)
packed = pack_gt_host(gts, pin_memory=False).to(dev)
crit = YoloDFLQFLoss(num_classes=nc)
def step():
    x.grad = None
    loss, parts = crit(x, packed, anchors, strides)
    loss.backward()
    return parts
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize(); t1 = time.perf_counter()
print('api step us', (t1 - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
