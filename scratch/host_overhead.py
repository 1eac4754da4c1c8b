import sys, time, torch, cProfile, pstats
sys.path.insert(0, '/root/repo')
from custom_yolo_implmentation_b200.model.losses import fused_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
n, nc = 128, 80
preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, 640, 100, 1236)
preds = preds.to(dev); anchors = anchors.to(dev); strides = strides.to(dev)
gt, off, counts = pack_gt([g.to(dev) for g in gts], dev)
def step(): return fused_loss(preds, gt, off, max(counts), anchors, strides, nc, 1.0, 1.5)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): o = step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host per call us', (t1 - t0) / 200 * 1e6, 'total per call us', (t2 - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): o = step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
# CUDA graph replay
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): step()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        o = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5): g.replay()
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
print('graph replay ms/step', e0.elapsed_time(e1) / 200, 'loss', o[0][0].item())
