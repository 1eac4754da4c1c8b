import ctypes, sys, torch
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from custom_yolo_implmentation_b200 import _cabi
from custom_yolo_implmentation_b200.model import losses as P
from custom_yolo_implmentation_b200.utils import synthetic as syn
dev=torch.device('cuda:0')
preds, gts, anchors, strides = syn.make_loss_inputs(128, 80, 640, 100, 1236)
gt, off, counts = P.pack_gt([g.to(dev) for g in gts], dev); x=preds.to(dev); a=anchors.to(dev); s=strides.to(dev)
lib=ctypes.CDLL(ROOT+'/scratch/variants/libtal_clocks.so')
lib.yb_tal_workspace_bytes.restype=ctypes.c_size_t; lib.yb_tal_workspace_bytes.argtypes=[ctypes.c_int]*5
Pp,I=ctypes.c_void_p,ctypes.c_int
lib.yb_tal_assign.argtypes=[Pp,I,I,I,I,I,Pp,Pp,Pp,Pp,I,Pp,Pp,Pp,Pp,Pp,Pp,Pp,ctypes.c_size_t,Pp]
prm=_cabi.TalParams(10,0.5,6.0,1.5,1.0,1.5,0,0.75,2.0)
n,c,A=x.shape; G=gt.shape[0]; hint=P.build_grid_hint(a,s)
ws=torch.zeros(lib.yb_tal_workspace_bytes(n,A,G,0,10),dtype=torch.uint8,device=dev); stats=torch.empty(8,device=dev)
def call():
    rc=lib.yb_tal_assign(x.data_ptr(),0,n,80,16,A,a.data_ptr(),s.data_ptr(),gt.data_ptr(),off.data_ptr(),G,ctypes.byref(prm),ctypes.byref(hint),None,stats.data_ptr(),None,None,ws.data_ptr(),ws.numel(),None); assert rc==0
for _ in range(3): call()
torch.cuda.synchronize(); buf=(ctypes.c_ulonglong*16)(); lib.yb_tal_clocks(buf)
K=5
for _ in range(K): call()
torch.cuda.synchronize(); lib.yb_tal_clocks(buf)
names=['A setup+seed','A enumerate+filter+eval','A publish','(of which evaluate2)','B wait','B scalar','B per-bin rounds','C resolve unit']
tot=sum(buf[i] for i in (0,1,2,4,5,6,7))
for i,nm in enumerate(names): print(f'{nm:28s} {buf[i]/K/G/1.9e3:8.2f} us per GT   {100*buf[i]/tot:5.1f} %')
print('warp-time per GT us', tot/K/G/1.9e3, ' x GTs / warps(3552) =', tot/K/1.9e3/3552, 'us')
