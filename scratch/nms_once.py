"""cfg4 NMS twice (for ncu captures)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200.utils import synthetic as syn
from custom_yolo_implmentation_b200.utils import model_utils as U
dev = torch.device('cuda:0')
y = syn.make_nms_input(64, 80, 640, 2024).to(dev)
for _ in range(2):
    rows, count, _ = U.batched_nms_raw(y, 0.001, 0.7, 300, 80, False, None)
torch.cuda.synchronize()
print('ok', count[:4].tolist())
