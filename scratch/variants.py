"""Build loss.cu variants with different tuning macros and time each on the GPU (not product code)."""
import ctypes, os, subprocess, sys, itertools, json
import torch
sys.path.insert(0, '/root/repo')
ROOT = '/root/repo'
CSRC = os.path.join(ROOT, 'custom-yolo-implmentation_b200', 'csrc')
OUT = os.path.join(ROOT, 'scratch', 'variants')
VARIANTS = (json.load(open(os.path.join(ROOT, 'scratch', 'variants.json'))) if os.path.exists(os.path.join(ROOT, 'scratch', 'variants.json')) else None) or [
    {}, {"YB_CLS_THREADS": 256}, {"YB_CLS_UNROLL": 8}, {"YB_CLS_UNROLL": 2}, {"YB_CLS_MINBLOCKS": 8},
    {"YB_CLS_THREADS": 256, "YB_CLS_UNROLL": 2}, {"YB_ASSIGN_THREADS": 256}, {"YB_ASSIGN_THREADS": 64}, {"YB_ASSIGN_MINBLOCKS": 8},
]
def name(v): return "base" if not v else v["_name"] if "_name" in v else '_'.join(f"{k[3:].lower()}{val}" for k, val in sorted(v.items()))
def build_all():
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for v in VARIANTS:
        so = os.path.join(OUT, f'lib_{name(v)}.so')
        defs = [f'-D{k}={val}' for k, val in v.items() if not k.startswith('_')]
        cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC',
               '--expt-relaxed-constexpr', '-shared', '-o', so, os.path.join(CSRC, 'loss.cu'), os.path.join(CSRC, 'cabi.cu'), '-lcudart'] + defs
        procs.append((v, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for v, p in procs:
        out, _ = p.communicate()
        if p.returncode: print('BUILD FAIL', name(v), out.decode()[-400:])
def run_all():
    from custom_yolo_implmentation_b200.utils import synthetic as syn
    dev = torch.device('cuda:0')
    cfgs = {'cfg2': (128, 80, 640, 100, torch.float32), 'cfg5': (32, 80, 1280, 300, torch.bfloat16)}
    data = {}
    for cn, (n, nc, imgsz, gmax, dt) in cfgs.items():
        preds, gts, anchors, strides = syn.make_loss_inputs(n, nc, imgsz, gmax, 1236, dtype=dt)
        counts = [g.shape[0] for g in gts]
        off = torch.tensor([0] + list(itertools.accumulate(counts)), dtype=torch.int32, device=dev)
        data[cn] = dict(preds=preds.to(dev), gt=torch.cat(gts).to(dev), off=off, anc=anchors.float().to(dev), st=strides.float().to(dev),
                        hint=(__import__('custom_yolo_implmentation_b200.model.losses', fromlist=['x']).build_grid_hint(*syn.anchor_grid(imgsz)) if os.environ.get('YB_NO_HINT') is None else None), n=n, nc=nc, a=preds.shape[2], gt_total=sum(counts), gmax=max(counts), dt=1 if dt == torch.bfloat16 else 0)
    for v in VARIANTS:
        so = os.path.join(OUT, f'lib_{name(v)}.so')
        if not os.path.exists(so): continue
        lib = ctypes.CDLL(so)
        lib.yb_loss_workspace_bytes.restype = ctypes.c_size_t
        lib.yb_loss_workspace_bytes.argtypes = [ctypes.c_int] * 4
        P = ctypes.c_void_p
        lib.yb_loss_fwd_bwd.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_float, P, P, P, P, P, P, ctypes.c_size_t, ctypes.c_uint, P, P, P]
        line = [f'{name(v):28s}']
        for cn, d in data.items():
            ws = torch.zeros(lib.yb_loss_workspace_bytes(d['n'], d['a'], d['gt_total'], d['dt']), dtype=torch.uint8, device=dev)
            grad = torch.empty_like(d['preds']); out = torch.empty(8, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            def call():
                rc = lib.yb_loss_fwd_bwd(d['preds'].data_ptr(), d['dt'], d['n'], d['nc'], 16, d['a'], d['anc'].data_ptr(), d['st'].data_ptr(), d['gt'].data_ptr(),
                                         d['off'].data_ptr(), d['gt_total'], d['gmax'], 1.0, 1.5, grad.data_ptr(), out.data_ptr(), None, None, None, ws.data_ptr(), ws.numel(), 8, (ctypes.byref(d['hint']) if d['hint'] is not None else None), None, st)
                assert rc == 0, lib.yb_last_error()
            for _ in range(3): call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200): call()
            e1.record(); torch.cuda.synchronize()
            total = e0.elapsed_time(e1) / 200
            tk = ws[:64].view(torch.int32).tolist(); line.append(f'{cn}: step {total*1e3:6.1f} us  loss {out[0].item():.6f} waits {tk[8]} kcyc {tk[9]}')
        print(' | '.join(line), flush=True)
if __name__ == '__main__':
    if sys.argv[1] == 'build': build_all()
    else: run_all()
