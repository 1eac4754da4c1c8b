"""NMS timing on the fall-back paths: agnostic (single-CTA sweep), few candidates, high conf."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.utils import model_utils as U
dev = torch.device('cuda:0')
y = syn.make_nms_input(64, 80, 640, 2024).to(dev)
for name, kw in (("class-aware conf 0.001", dict(conf=0.001, agn=False)), ("agnostic conf 0.001", dict(conf=0.001, agn=True)),
                 ("class-aware conf 0.25", dict(conf=0.25, agn=False)), ("agnostic conf 0.25", dict(conf=0.25, agn=True))):
    f = lambda: U.batched_nms_raw(y, kw["conf"], 0.7, 300, 80, kw["agn"], None)
    for _ in range(3): rows, count, _ = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): rows, count, _ = f()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:26s} {e0.elapsed_time(e1)/20*1e3:9.1f} us/call  kept {count[:3].tolist()}')
for name, conf in (("multi_label conf 0.001", 0.001), ("multi_label conf 0.25", 0.25)):
    f = lambda: U.batched_nms_raw(y, conf, 0.7, 300, 80, False, None, multi_label=True)
    for _ in range(3): rows, count, _ = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): rows, count, _ = f()
    e1.record(); torch.cuda.synchronize()
    print(f'{name:26s} {e0.elapsed_time(e1)/10*1e3:9.1f} us/call  kept {count[:3].tolist()}')
