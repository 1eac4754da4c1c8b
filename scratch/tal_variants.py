"""Build tal.cu variants with tuning macros and time yb_tal_assign / yb_tal_loss (not product code).
usage: tal_variants.py build|run   (variants from scratch/tal_variants.json)"""
import ctypes, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
CSRC = os.path.join(ROOT, 'custom-yolo-implmentation_b200', 'csrc')
OUT = os.path.join(ROOT, 'scratch', 'variants')
VARIANTS = json.load(open(os.environ.get('YB_VARIANTS', os.path.join(ROOT, 'scratch', 'tal_variants.json'))))
def name(v): return v['_name'] if '_name' in v else 'base' if not v else '_'.join(f"{k[3:].lower()}{val}" for k, val in sorted(v.items()))
def build_all():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for v in VARIANTS:
        so = os.path.join(OUT, f'libtal_{name(v)}.so')
        cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
               '--expt-relaxed-constexpr', '-shared', '-o', so, os.path.join(ROOT, v['_src']) if '_src' in v else os.path.join(CSRC, 'tal.cu'), os.path.join(CSRC, 'peer.cu'), os.path.join(CSRC, 'cabi.cu'), '-lcudart', '-I', CSRC] + [f'-D{k}={val}' for k, val in v.items() if not k.startswith('_')]
        procs.append((v, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for v, p in procs:
        out, _ = p.communicate()
        if p.returncode: print('BUILD FAIL', name(v), out.decode()[-600:])
def run_all():
    import torch
    from custom_yolo_implmentation_b200 import _cabi
    from custom_yolo_implmentation_b200.model import losses as P
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    from test_gpu_tal import make_inputs
    dev = torch.device('cuda:0')
    data = {}
    for cn in ('bench', 'spread', 'bf16'):
        if cn == 'bench': preds, gts, anchors, strides = syn.make_loss_inputs(128, 80, 640, 100, 1236)
        else: preds, gts, anchors, strides = make_inputs(128, 80, 640, 100, 51)
        gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev)
        dt, code = (torch.bfloat16, 1) if cn == 'bf16' else (torch.float32, 0)
        data[cn] = (preds.to(dev, dt), gt, off, anchors.to(dev), strides.to(dev), code)
    for v in VARIANTS:
        so = os.path.join(OUT, f'libtal_{name(v)}.so')
        if not os.path.exists(so): continue
        lib = ctypes.CDLL(so)
        lib.yb_tal_workspace_bytes.restype = ctypes.c_size_t
        lib.yb_tal_workspace_bytes.argtypes = [ctypes.c_int] * 5
        Pp, I = ctypes.c_void_p, ctypes.c_int
        lib.yb_tal_assign.argtypes = [Pp, I, I, I, I, I, Pp, Pp, Pp, Pp, I, Pp, Pp, Pp, Pp, Pp, Pp, Pp, ctypes.c_size_t, Pp]
        lib.yb_tal_loss.argtypes = [Pp, I, I, I, I, I, I, Pp, Pp, Pp, Pp, Pp, Pp, ctypes.c_size_t, Pp]
        lib.yb_last_error.restype = ctypes.c_char_p
        prm = _cabi.TalParams(10, 0.5, 6.0, 1.5, 1.0, 1.5, 0, 0.75, 2.0, 1)
        line = [f'{name(v):24s}']
        for cn, (x, gt, off, a, s, code) in data.items():
            n, c, A = x.shape; G = gt.shape[0]
            hint = P.build_grid_hint(a, s) if os.environ.get('YB_NO_HINT') is None else None
            ws = torch.zeros(lib.yb_tal_workspace_bytes(n, A, G, code, 10), dtype=torch.uint8, device=dev)
            stats = torch.empty(8, device=dev); out = torch.empty(8, device=dev); grad = torch.empty_like(x)
            st = torch.cuda.current_stream().cuda_stream
            def assign():
                rc = lib.yb_tal_assign(x.data_ptr(), code, n, 80, 16, A, a.data_ptr(), s.data_ptr(), gt.data_ptr(), off.data_ptr(), G, ctypes.byref(prm), ctypes.byref(hint) if hint is not None else None, None, stats.data_ptr(), None, None, ws.data_ptr(), ws.numel(), st)
                assert rc == 0, lib.yb_last_error()
            def loss():
                rc = lib.yb_tal_loss(x.data_ptr(), code, n, 80, 16, A, G, ctypes.byref(prm), stats.data_ptr(), None, grad.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), st)
                assert rc == 0, lib.yb_last_error()
            for _ in range(3): assign(); loss()
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ta = tl = 0
            for _ in range(30):
                ev[0].record(); assign(); ev[1].record(); loss(); ev[2].record(); torch.cuda.synchronize()
                ta += ev[0].elapsed_time(ev[1]); tl += ev[1].elapsed_time(ev[2])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100): assign(); loss()
            e1.record(); torch.cuda.synchronize()
            line.append(f'{cn}: assign {ta/30*1e3:6.1f} loss {tl/30*1e3:6.1f} step {e0.elapsed_time(e1)/100*1e3:6.1f} us  L={out[0].item():.5f} gsum={grad.float().abs().sum().item():.4f}')
        print(' | '.join(line), flush=True)
if __name__ == '__main__':
    build_all() if sys.argv[1] == 'build' else run_all()
