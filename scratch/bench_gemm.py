"""Microbenchmark of edadm_qgemm_i8 on the dominant QuantModule shapes (event-timed, rotating buffers > L2)."""
import sys, os, json
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
ops.w4_storage = bool(int(os.environ.get('W4','1')))
dev=torch.device('cuda:0')
# (name, B, C, H, N, k)  conv with pad=k//2 ; k=0 -> linear with M=B
SHAPES=[("church up c576->192 32x32",100,576,32,192,3),("church up c384->192 32x32",100,384,32,192,3),("church up c1152->384 16x16",100,1152,16,384,3),("church 32x32 c192 3x3",100,192,32,192,3),("church 16x16 c384 3x3",100,384,16,384,3),("church 8x8 c384 3x3",100,384,8,384,3),
        ("church 4x4 c768 3x3",100,768,4,768,3),("church qkv 1x1 T1024",100,192,32,576,1),("imagenet 64x64 c192 3x3",32,192,64,192,3),
        ("imagenet 32x32 c384 3x3",32,384,32,384,3),("imagenet 16x16 c576",32,576,16,576,3),("imagenet geglu lin",32*1024,384,0,3072,0),
        ("imagenet lin 384",32*1024,384,0,384,0),("bedroom 64x64 c224",8,224,64,224,3)]
only=os.environ.get("ONLY")
reps=int(os.environ.get("REPS","20"))
res=[]
for name,B,C,H,N,k in SHAPES:
    if only and only not in name: continue
    torch.manual_seed(0)
    d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev)
    aq=ops.ActQuant(d,z,256)
    if k:
        w=torch.randn(N,C,k,k,device=dev)*0.05
        x=torch.randn(B,C,H,H,device=dev)
        M=B*H*H; K=C*k*k
    else:
        w=torch.randn(N,C,device=dev)*0.05
        x=torch.randn(B,C,device=dev)
        M=B; K=C
    dw=(w.flatten(1).abs().amax(1)/7.5).reshape(-1,*([1]*(w.dim()-1))); zw=torch.full_like(dw,8.)
    pw=ops.pack_weight(w,dw,zw,16)
    nbuf=4
    if k:
        qs=[ops.act_quant_nhwc(x,aq,k//2)[0] for _ in range(nbuf)]
        outs=[torch.empty(B,N,H,H,device=dev) for _ in range(nbuf)]
        hw=H*H
    else:
        qs=[ops.act_quant_rows(x,aq)[0] for _ in range(nbuf)]
        outs=[torch.empty(B,N,device=dev) for _ in range(nbuf)]
        hw=1
    for i in range(3): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw)
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*1e3/reps
    tops=2*M*N*K/us/1e6
    line=f"{name:28s} M={M:7d} N={N:5d} K={K:6d}  {us:8.1f} us  {tops:7.1f} TOP/s"
    ress=[torch.randn_like(outs[0]) for _ in range(nbuf)]
    for i in range(3): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,residual=ress[i%nbuf])
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,residual=ress[i%nbuf])
    e1.record(); torch.cuda.synchronize()
    us3=e0.elapsed_time(e1)*1e3/reps
    e0.record()
    for i in range(reps): torch.add(outs[i%nbuf],ress[i%nbuf],out=outs[(i+1)%nbuf])
    e1.record(); torch.cuda.synchronize()
    line+=f"  +res {us3:7.1f} us (add alone {e0.elapsed_time(e1)*1e3/reps:6.1f})"
    # library yardstick: torch._int_mm on the im2col'd problem size
    try:
        if os.environ.get("NOLIB"): raise RuntimeError("skipped")
        a=torch.randint(-128,127,(M,K),device=dev,dtype=torch.int8); b=torch.randint(-8,7,(K,N),device=dev,dtype=torch.int8)
        for i in range(2): torch._int_mm(a,b)
        torch.cuda.synchronize(); e0.record()
        for i in range(10): torch._int_mm(a,b)
        e1.record(); torch.cuda.synchronize()
        us2=e0.elapsed_time(e1)*1e3/10
        line+=f"   | cublasLt int8 (explicit GEMM, int32 out) {us2:8.1f} us {2*M*N*K/us2/1e6:7.1f} TOP/s"
    except Exception as e:
        line+=f"   | _int_mm n/a ({str(e)[:40]})"
    print(line, flush=True)
