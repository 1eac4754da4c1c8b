import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import loss_oracle as L
from custom_yolo_implmentation_b200.utils import synthetic as syn
from test_gpu_loss import run_cuda_trace, idx_agreement
dev=torch.device('cuda:0')
n,nc,imgsz,gmax,seed,conflict=3,80,640,100,1236,0.05
preds,gts,anchors,strides=syn.make_loss_inputs(n,nc,imgsz,gmax,seed,conflict_frac=conflict)
out,grad,idx,iou,dfl_img,cls_img=run_cuda_trace(preds,gts,anchors,strides,nc,dev)
ora=L.loss_forward_backward(preds,gts,anchors,strides,nc)
tot,bad=idx_agreement(idx,ora); print('tot',tot,'bad',bad)
if bad: ora=L.loss_forward_backward(preds,gts,anchors,strides,nc,forced_idx=idx)
print('dfl gpu',dfl_img.tolist(),'ora',ora.dfl_per_image.tolist())
print('cls gpu',cls_img.tolist(),'ora',ora.cls_per_image.tolist())
for b in range(n):
    if len(iou[b]):
        d=(iou[b]-ora.iou[b]).abs(); j=d.argmax()
        print(b,'iou range',iou[b].min().item(),iou[b].max().item(),'maxdiff',d.max().item(),'at',j.item(),iou[b][j].item(),ora.iou[b][j].item(), 'dups', len(idx[b])-len(idx[b].unique()))
g=(grad-ora.grad).abs(); print('grad err',g.max().item()/ora.grad.abs().max().item())
