"""Timeline of the GEMM's warp roles (clock64 stamps via edadm_debug_set_gemm_trace) on three conv shapes."""
import sys, os, ctypes
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops, native
ops.w4_storage = False
dev=torch.device('cuda:0')
h = native.load_library()
h.edadm_debug_set_gemm_trace.argtypes=[ctypes.c_void_p]
for name,B,C,H,N,k in [("imagenet 64x64 c192 3x3",32,192,64,192,3),("c256 64x64 3x3",32,256,64,192,3),("c128 64x64 3x3",32,128,64,192,3),("church 16x16 c384 3x3",100,384,16,384,3),("church up c1152->384",100,1152,16,384,3)]:
    d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev)
    aq=ops.ActQuant(d,z,256)
    w=torch.randn(N,C,k,k,device=dev)*0.05
    x=torch.randn(B,C,H,H,device=dev)
    dw=(w.flatten(1).abs().amax(1)/7.5).reshape(-1,1,1,1); zw=torch.full_like(dw,8.)
    pw=ops.pack_weight(w,dw,zw,16)
    q=ops.act_quant_nhwc(x,aq,1,cp=int(os.environ.get('CPACT','0')))[0]
    out=torch.empty(B,N,H,H,device=dev)
    for _ in range(3): ops.qgemm_i8(q,pw,d,z,out,H*H)
    trace=torch.zeros(148*16*8,dtype=torch.int64,device=dev)
    h.edadm_debug_set_gemm_trace(trace.data_ptr())
    ops.qgemm_i8(q,pw,d,z,out,H*H)
    torch.cuda.synchronize()
    h.edadm_debug_set_gemm_trace(None)
    t=trace.reshape(148,16,8).cpu()
    print("==",name)
    for cta in (0,):
        base=int(t[cta,0,5])
        for ti in range(8):
            r=t[cta,ti]
            if int(r[0])==0: break
            print(f" cta {cta} tile {ti}: prod {int(r[5])-base:7d}..{int(r[6])-base:7d} | mma start {int(r[0])-base:7d} first-full {int(r[1])-base:7d} end {int(r[2])-base:7d} (dur {int(r[2])-int(r[0]):6d}) | epi wake {int(r[3])-base:7d} done {int(r[4])-base:7d} (dur {int(r[4])-int(r[3]):6d})")
