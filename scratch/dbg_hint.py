import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200 import _cabi
from test_gpu_tal import make_inputs, run_cuda
dev = torch.device('cuda:0')
preds, gts, anchors, strides = make_inputs(3, 80, 640, 60, 36)
ref = run_cuda(preds, gts, anchors, strides, 80, dev)
for trial in range(4):
    for h in (None, "auto"):
        got = run_cuda(preds, gts, anchors, strides, 80, dev, grid_hint=h)
        d0 = (got[0][:6] - ref[0][:6]).abs().tolist()
        dg = (got[1] - ref[1]).abs()
        print(trial, h, 'out diff', d0, 'grad maxdiff', dg.max().item(), 'n diff', int((dg > 0).sum()), 'trace eq', [torch.equal(a, b) for a, b in zip(got[2:], ref[2:])])
