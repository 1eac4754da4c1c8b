"""Time cfg4 NMS (64 x 8400 x 80, conf 0.001, iou 0.7) with CUDA events."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.utils import model_utils as U
dev = torch.device('cuda:0')
for n in (64, 16, 8, 32, 128):
    y = syn.make_nms_input(n, 80, 640, 2024).to(dev)
    for _ in range(5): rows, count, _ = U.batched_nms_raw(y, 0.001, 0.7, 300, 80, False, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): rows, count, _ = U.batched_nms_raw(y, 0.001, 0.7, 300, 80, False, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f'N={n}: {ms*1e3:.1f} us/call  {n/ms*1e3:.0f} img/s  kept {count[:3].tolist()}')
