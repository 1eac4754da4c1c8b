"""Phase times of tal_gt_kernel's warps (YB_TAL_TRACE build): start, end of the selection loop, of the terms loop, of the
target-score loop; relative to the first warp's start."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = '/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
CSRC = os.path.join(ROOT, 'custom-yolo-implmentation_b200', 'csrc')
SO = os.path.join(ROOT, 'scratch', 'variants', 'libtal_trace.so')
if sys.argv[1] == 'build':
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
                           '--expt-relaxed-constexpr', '-shared', '-o', SO, os.path.join(CSRC, 'tal.cu'), os.path.join(CSRC, 'peer.cu'), os.path.join(CSRC, 'cabi.cu'), '-lcudart', '-DYB_TAL_TRACE=1'])
    sys.exit(0)
from custom_yolo_implmentation_b200 import _cabi
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
lib = ctypes.CDLL(SO)
lib.yb_tal_workspace_bytes.restype = ctypes.c_size_t
lib.yb_tal_workspace_bytes.argtypes = [ctypes.c_int] * 5
Pp, I = ctypes.c_void_p, ctypes.c_int
lib.yb_tal_assign.argtypes = [Pp, I, I, I, I, I, Pp, Pp, Pp, Pp, I, Pp, Pp, Pp, Pp, Pp, Pp, Pp, ctypes.c_size_t, Pp]
lib.yb_tal_loss.argtypes = [Pp, I, I, I, I, I, I, Pp, Pp, Pp, Pp, Pp, Pp, ctypes.c_size_t, Pp]
lib.yb_tal_trace_dump.argtypes = [Pp, Pp]
preds, gts, anchors, strides = syn.make_loss_inputs(128, 80, 640, 100, 1236)
gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev)
x = preds.to(dev); a = anchors.to(dev); s = strides.to(dev)
n, c, A = x.shape; G = gt.shape[0]
prm = _cabi.TalParams(10, 0.5, 6.0, 1.5, 1.0, 1.5, 0, 0.75, 2.0, 1)
hint = P.build_grid_hint(a, s)
ws = torch.zeros(lib.yb_tal_workspace_bytes(n, A, G, 0, 10), dtype=torch.uint8, device=dev)
stats = torch.empty(8, device=dev); out = torch.empty(8, device=dev); grad = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    assert lib.yb_tal_assign(x.data_ptr(), 0, n, 80, 16, A, a.data_ptr(), s.data_ptr(), gt.data_ptr(), off.data_ptr(), G, ctypes.byref(prm), ctypes.byref(hint), None, stats.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st) == 0
    assert lib.yb_tal_loss(x.data_ptr(), 0, n, 80, 16, A, G, ctypes.byref(prm), stats.data_ptr(), None, grad.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st) == 0
torch.cuda.synchronize()
buf = np.zeros((1 << 13, 4), dtype=np.uint64); buf2 = np.zeros(1 << 13, dtype=np.uint64)
assert lib.yb_tal_trace_dump(buf.ctypes.data, buf2.ctypes.data) == 0
m = buf[:, 0] > 0
buf, buf2 = buf[m], buf2[m]
t0 = buf[:, 0].min()
rel = (buf - t0).astype(np.float64) / 1e3
end2 = (buf2 - t0).astype(np.float64) / 1e3
print(f'{len(buf)} warps')
for i, name in enumerate(('start', 'selection done', 'terms done', 'target scores done')):
    v = rel[:, i]
    print(f'  {name:20s} min {v.min():6.1f} p10 {np.percentile(v, 10):6.1f} p50 {np.median(v):6.1f} p90 {np.percentile(v, 90):6.1f} max {v.max():6.1f} us')
print(f'  after griddep wait   min {end2.min():6.1f} p50 {np.median(end2):6.1f} max {end2.max():6.1f} us')
