"""Timeline of fused_main_kernel's CTAs (YB_LOSS_TRACE build): role, SM, start, end -> occupancy over time per role."""
import ctypes, os, subprocess, sys, itertools, json
import numpy as np, torch
sys.path.insert(0, '/root/repo')
ROOT = '/root/repo'
CSRC = os.path.join(ROOT, 'custom-yolo-implmentation_b200', 'csrc')
SO = os.path.join(ROOT, 'scratch', 'variants', 'lib_trace.so')
if sys.argv[1] == 'build':
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    extra = sys.argv[2:]
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
                           '--expt-relaxed-constexpr', '-shared', '-o', SO, os.path.join(CSRC, 'loss.cu'), os.path.join(CSRC, 'cabi.cu'),
                           '-lcudart', '-DYB_LOSS_TRACE=1'] + extra)
    sys.exit(0)
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.model.losses import build_grid_hint
dev = torch.device('cuda:0')
lib = ctypes.CDLL(SO)
lib.yb_loss_workspace_bytes.restype = ctypes.c_size_t
lib.yb_loss_workspace_bytes.argtypes = [ctypes.c_int] * 4
P = ctypes.c_void_p
lib.yb_loss_fwd_bwd.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.c_int,
                                ctypes.c_float, ctypes.c_float, P, P, P, P, P, P, ctypes.c_size_t, ctypes.c_uint, P, P, P]
lib.yb_trace_dump.argtypes = [P]
names = {1: 'box-fine', 2: 'box-coarse', 3: 'class', 11: 'probe', 12: 'match', 13: 'reducer', 0: 'none'}
for cn, (n, nc, imgsz, gmax, dt) in {'cfg2': (128, 80, 640, 100, torch.float32), 'cfg5': (32, 80, 1280, 300, torch.bfloat16)}.items():
    preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, 1236, dtype=dt)
    counts = [g.shape[0] for g in gts]
    off = torch.tensor([0] + list(itertools.accumulate(counts)), dtype=torch.int32, device=dev)
    preds = preds.to(dev); gt = torch.cat(gts).to(dev); anc = anchors.float().to(dev); st = strides.float().to(dev)
    hint = build_grid_hint(*syn.anchor_grid(imgsz))
    a = preds.shape[2]; gt_total = sum(counts); dtc = 1 if dt == torch.bfloat16 else 0
    ws = torch.zeros(lib.yb_loss_workspace_bytes(n, a, gt_total, dtc), dtype=torch.uint8, device=dev)
    grad = torch.empty_like(preds); out = torch.empty(8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        rc = lib.yb_loss_fwd_bwd(preds.data_ptr(), dtc, n, nc, 16, a, anc.data_ptr(), st.data_ptr(), gt.data_ptr(), off.data_ptr(), gt_total,
                                 max(counts), 1.0, 1.5, grad.data_ptr(), out.data_ptr(), None, None, None, ws.data_ptr(), ws.numel(), 8,
                                 ctypes.byref(hint), None, stream)
        assert rc == 0
    torch.cuda.synchronize()
    buf = np.zeros((1 << 16, 4), dtype=np.uint64)
    assert lib.yb_trace_dump(buf.ctypes.data) == 0
    red = buf[-1].copy(); buf = buf[:-1]
    buf = buf[buf[:, 3] > 0]
    buf = buf[buf[:, 1] + np.uint64(2_000_000) > buf[:, 1].max()]      # this launch only (entries of larger earlier grids linger)
    t0 = buf[:, 0].min(); t1 = buf[:, 1].max()
    s = (buf[:, 0] - t0).astype(np.float64) / 1e3; e = (buf[:, 1] - t0).astype(np.float64) / 1e3; role = buf[:, 3].astype(int)
    print(f'== {cn}: {len(buf)} CTAs, span {float(t1 - t0) / 1e3:.1f} us; reducer: waited until {(float(red[0]) - float(t0)) / 1e3:.1f}, reduced by {(float(red[1]) - float(t0)) / 1e3:.1f}')
    for r in sorted(set(role)):
        m = role == r
        d = e[m] - s[m]
        print(f'  {names.get(r, r):10s} n={m.sum():6d} dur mean {d.mean():6.1f} p50 {np.median(d):6.1f} p95 {np.percentile(d, 95):6.1f} max {d.max():6.1f} us | first start {s[m].min():6.1f} last start {s[m].max():6.1f} last end {e[m].max():6.1f}')
    # resident CTAs over time, by role, sampled every 10 us
    T = np.arange(0, float(t1 - t0) / 1e3 + 5, 5.0)
    print('  t(us)   ' + ' '.join(f'{names.get(r, r)[:9]:>9s}' for r in sorted(set(role))) + '   total')
    for t in T:
        row = [int(((s <= t) & (e > t) & (role == r)).sum()) for r in sorted(set(role))]
        print(f'  {t:6.0f}   ' + ' '.join(f'{x:9d}' for x in row) + f'   {sum(row):5d}')
