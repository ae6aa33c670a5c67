"""Role timeline (clock64 stamps) of the second-generation GEMM on the GEGLU projection and the q/k code GEMM"""
import sys, os, ctypes
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops, native
dev=torch.device('cuda:0')
h = native.load_library()
h.edadm_debug_set_gemm_trace.argtypes=[ctypes.c_void_p]
for (M,N,K,geglu) in [(131072,3072,384,True),(131072,384,384,False)]:
    g=torch.Generator().manual_seed(0)
    w=(torch.randn(N,K,generator=g)*0.05).to(dev)
    dw=(w.abs().amax(1)*2/15).clamp_min(1e-8); zw=torch.full((N,),8.0,device=dev)
    pw=ops.pack_weight(w, dw, zw, 16)
    q=torch.randint(0,256,(M,K),dtype=torch.uint8,generator=g).to(dev)
    da=torch.tensor([0.02],device=dev); za=torch.tensor([128.],device=dev)
    cons=(torch.tensor([0.05],device=dev), torch.tensor([128.],device=dev), 256)
    for _ in range(3): ops.qgemm_i8_codes(q, pw, da, za, cons, geglu=geglu)
    trace=torch.zeros(148*16*8,dtype=torch.int64,device=dev)
    h.edadm_debug_set_gemm_trace(trace.data_ptr())
    ops.qgemm_i8_codes(q, pw, da, za, cons, geglu=geglu)
    torch.cuda.synchronize()
    h.edadm_debug_set_gemm_trace(None)
    t=trace.reshape(148,16,8).cpu()
    print("==",(M,N,K,geglu))
    for cta in (0,77):
        base=int(t[cta,0,5])
        for ti in range(2,9):
            r=[int(v)-base for v in t[cta,ti]]
            print(f" cta {cta} tile {ti}: prod {r[5]:7d}..{r[6]:7d} | mma free {r[0]:7d} first-full {r[1]:7d} committed {r[2]:7d} | epi acc-ready {r[3]:7d} tmem-read {r[7]:7d} done {r[4]:7d} (epi {r[4]-r[3]:6d}, tile period {int(t[cta,ti,4])-int(t[cta,ti-1,4]):6d})")
