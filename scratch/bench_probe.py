import sys, time, torch
sys.path.insert(0,'/root/repo')
from custom_yolo_implmentation_b200.model.losses import fused_loss, pack_gt
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev=torch.device('cuda:0')
n,nc=128,80
preds,gts,anchors,strides=syn.make_loss_inputs(n,nc,640,100,1236)
preds=preds.to(dev); a=anchors.to(dev); s=strides.to(dev)
gt,off,counts=pack_gt([g.to(dev) for g in gts],dev)
def step(): return fused_loss(preds,gt,off,max(counts),a,s,nc,1.0,1.5)
for _ in range(10): step()
torch.cuda.synchronize()
e=[torch.cuda.Event(enable_timing=True) for _ in range(60)]
for trial in range(3):
    torch.cuda.synchronize()
    t0=time.perf_counter()
    e[0].record()
    for i in range(50):
        o=step(); e[i+1].record()
    t1=time.perf_counter()
    torch.cuda.synchronize()
    per=[e[i].elapsed_time(e[i+1]) for i in range(50)]
    print('trial',trial,'host enqueue ms',(t1-t0)*1e3,'total ms',e[0].elapsed_time(e[50]),'first 6 steps ms',[round(x,3) for x in per[:6]],'last',round(per[-1],3))
import pynvml
pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
torch.cuda.synchronize(); e[0].record()
for i in range(50):
    o=step(); e[i+1].record()
torch.cuda.synchronize()
print('after nvml init: total ms',e[0].elapsed_time(e[50]))
