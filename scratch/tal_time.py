import sys, torch, time, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200 import _cabi
from test_gpu_tal import make_inputs
dev=torch.device('cuda:0')
preds,gts,anchors,strides=make_inputs(128,80,640,100,51)
gt,off,counts=P.pack_gt([g.to(dev) for g in gts],dev); x=preds.to(dev); a=anchors.to(dev); s=strides.to(dev)
lib=_cabi.lib(); n,c,A=x.shape; G=gt.shape[0]
ws=torch.empty(lib.yb_tal_workspace_bytes(n,A,G,0,10),dtype=torch.uint8,device=dev)
stats=torch.empty(8,device=dev); out=torch.empty(8,device=dev); grad=torch.empty_like(x)
st=torch.cuda.current_stream().cuda_stream
def assign(): 
    rc=lib.yb_tal_assign(x.data_ptr(),0,n,80,16,A,a.data_ptr(),s.data_ptr(),gt.data_ptr(),off.data_ptr(),G,10,0.5,6.0,stats.data_ptr(),None,None,ws.data_ptr(),ws.numel(),st); assert rc==0
def loss():
    rc=lib.yb_tal_loss(x.data_ptr(),0,n,80,16,A,a.data_ptr(),s.data_ptr(),gt.data_ptr(),off.data_ptr(),G,10,stats.data_ptr(),1.5,1.0,1.5,grad.data_ptr(),out.data_ptr(),ws.data_ptr(),ws.numel(),st); assert rc==0
for _ in range(3): assign(); loss()
torch.cuda.synchronize()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
ta=tl=0; t0=time.perf_counter()
for _ in range(50):
    ev[0].record(); assign(); ev[1].record(); loss(); ev[2].record(); torch.cuda.synchronize()
    ta+=ev[0].elapsed_time(ev[1]); tl+=ev[1].elapsed_time(ev[2])
print('assign ms',ta/50,'loss ms',tl/50)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(100): assign(); loss()
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print('host per step us',(t1-t0)/100*1e6,'total per step us',(t2-t0)/100*1e6)
import cProfile, pstats
def step(): return P.fused_tal_loss(x,gt,off,a,s,80,1.5,1.0,1.5)
for _ in range(3): step()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(100): o=step()
t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
print('wrapper: host per step us',(t1-t0)/100*1e6,'total per step us',(t2-t0)/100*1e6)
pr=cProfile.Profile(); pr.enable()
for _ in range(50): o=step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
