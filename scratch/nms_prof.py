import ctypes, sys, torch
sys.path.insert(0,'/root/repo')
from custom_yolo_implmentation_b200.utils import synthetic as syn
lib=ctypes.CDLL('/root/repo/scratch/variants/lib_nmsprof.so')
P=ctypes.c_void_p
lib.yb_nms_workspace_bytes.restype=ctypes.c_size_t; lib.yb_nms_workspace_bytes.argtypes=[ctypes.c_int]*2
lib.yb_nms.argtypes=[P,ctypes.c_int,ctypes.c_int,ctypes.c_int,ctypes.c_float,ctypes.c_double,ctypes.c_int,ctypes.c_int,P,ctypes.c_int,P,P,P,P,ctypes.c_size_t,P]
dev=torch.device('cuda:0')
y=syn.make_nms_input(64,80,640,2024).to(dev)
ws=torch.empty(lib.yb_nms_workspace_bytes(64,8400),dtype=torch.uint8,device=dev)
rows=torch.empty(64,300,6,device=dev); cnt=torch.empty(64,dtype=torch.int32,device=dev)
for _ in range(3):
    rc=lib.yb_nms(y.data_ptr(),64,80,8400,0.001,0.7,300,0,None,0,rows.data_ptr(),cnt.data_ptr(),None,ws.data_ptr(),ws.numel(),None); assert rc==0
torch.cuda.synchronize()
buf=(ctypes.c_longlong*16)(); lib.yb_nms_profile_read(buf)
names=['hist+scan','scatter','ranksort','gather','greedy','sel-setup+radix','collect','bitonic','emit']
for i,nm in enumerate(names): print(f'{nm:18s} {(buf[i+1]-buf[i])/1.9e3:8.1f} us')
print('kept',cnt[:4].tolist())
wb=(ctypes.c_longlong*128)(); lib.yb_nms_profile_read_warps(wb)
for w in range(0,32,4): print('warp',w,'resolve us',wb[w*4]/1.9e3/3,'class-loop us',wb[w*4+1]/1.9e3/3,'members',wb[w*4+2]/3,'classes',wb[w*4+3]/3)
