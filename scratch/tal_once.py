"""Two TAL steps at cfg2 (for ncu captures)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from custom_yolo_implmentation_b200.model import losses as P
from test_gpu_tal import make_inputs
dev = torch.device('cuda:0')
dtype = torch.bfloat16 if len(sys.argv) > 1 and sys.argv[1] == 'bf16' else torch.float32
preds, gts, anchors, strides = make_inputs(128, 80, 640, 100, 51)
gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev)
x = preds.to(dev, dtype); a = anchors.to(dev); s = strides.to(dev)
for _ in range(2):
    o = P.fused_tal_loss(x, gt, off, a, s, 80, 1.5, 1.0, 1.5)
torch.cuda.synchronize()
print('ok', float(o[0][0]) if isinstance(o, tuple) else o)
