"""Time cfg4 NMS for the .so variants given on the command line (not product code)."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev = torch.device('cuda:0')
y = syn.make_nms_input(64, 80, 640, 2024).to(dev)
P = ctypes.c_void_p
ref = None
for so in sys.argv[1:]:
    lib = ctypes.CDLL(so)
    lib.yb_nms_workspace_bytes.restype = ctypes.c_size_t; lib.yb_nms_workspace_bytes.argtypes = [ctypes.c_int] * 2
    lib.yb_nms.argtypes = [P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_int, ctypes.c_int, P, ctypes.c_int, P, P, P, P, ctypes.c_size_t, P]
    ws = torch.empty(lib.yb_nms_workspace_bytes(64, 8400), dtype=torch.uint8, device=dev)
    rows = torch.zeros(64, 300, 6, device=dev); cnt = torch.empty(64, dtype=torch.int32, device=dev)
    call = lambda: lib.yb_nms(y.data_ptr(), 64, 80, 8400, 0.001, 0.7, 300, 0, None, 0, rows.data_ptr(), cnt.data_ptr(), None, ws.data_ptr(), ws.numel(), None)
    for _ in range(5): assert call() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): call()
    e1.record(); torch.cuda.synchronize()
    if ref is None: ref = rows.clone()
    print(f"{so.split('/')[-1]:28s} {e0.elapsed_time(e1) * 10:.1f} us/call  rows equal to first: {torch.equal(rows, ref)}", flush=True)
