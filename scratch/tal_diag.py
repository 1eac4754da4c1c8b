import sys, torch, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import tal_oracle as T
from custom_yolo_implmentation_b200.model import losses as P
from test_gpu_tal import run_cuda, make_inputs
dev=torch.device('cuda:0')
for (n,nc,imgsz,gmax,seed,topk) in [(3,80,640,50,31,10),(2,20,320,120,32,10),(2,3,96,6,33,4),(2,80,640,30,34,13)]:
    preds,gts,anchors,strides=make_inputs(n,nc,imgsz,gmax,seed)
    out,grad,asg,tsc,stats=run_cuda(preds,gts,anchors,strides,nc,dev,topk=topk)
    ora=T.tal_forward_backward(preds,gts,anchors,strides,nc,topk=topk)
    print('case',seed,'diff',int((asg!=ora.assigned_gt).sum()),'fg',ora.num_fg,'loss',out[:4].tolist(),[ora.total.item(),ora.box.item(),ora.cls.item(),ora.dfl.item()],'gerr',(grad-ora.grad).abs().max().item()/ora.grad.abs().max().item())
# timing at cfg2
preds,gts,anchors,strides=make_inputs(128,80,640,100,51)
gt,off,counts=P.pack_gt([g.to(dev) for g in gts],dev); x=preds.to(dev); a=anchors.to(dev); s=strides.to(dev)
for _ in range(5): P.fused_tal_loss(x,gt,off,a,s,80,1.5,1.0,1.5)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): o=P.fused_tal_loss(x,gt,off,a,s,80,1.5,1.0,1.5)
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/100
print('TAL cfg2 ms/step',ms,'img/s',128/ms*1e3,'loss',o[0][:6].tolist())
