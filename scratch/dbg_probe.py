import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200 import _cabi
from custom_yolo_implmentation_b200.utils import synthetic as syn
from test_gpu_loss import run_cuda_trace
dev = torch.device('cuda:0')
preds, gts, anchors, strides = syn.make_loss_inputs(3, 80, 1280, 300, 61)
ref = run_cuda_trace(preds, gts, anchors, strides, 80, dev, flags=_cabi.YB_LOSS_NO_PRUNE, grid_hint=None)
for kw in (dict(), dict(grid_hint=None)):
    got = run_cuda_trace(preds, gts, anchors, strides, 80, dev, **kw)
    print(kw, 'out eq', torch.equal(got[0], ref[0]), (got[0]-ref[0]).abs().max().item())
    for b in range(3):
        d = (got[2][b] != ref[2][b]).nonzero().flatten().tolist()
        print(' image', b, 'idx diffs', len(d), d[:5], [(int(got[2][b][i]), int(ref[2][b][i])) for i in d[:5]])
# the workspace's bound array: recompute true min distances on the host for image 0
from oracle import loss_oracle as L
