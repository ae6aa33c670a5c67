"""Correctness (vs an exact fp64 convolution of the integer codes) + timing of edadm_qgemm_i8 on the dominant conv shapes."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch, torch.nn.functional as F
from edadm import ops
ops.w4_storage = False
dev=torch.device('cuda:0')
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
SHAPES=[("in c192 64x64",32,192,64,192,3),("in c384 32x32",32,384,32,384,3),("in c576 16x16",32,576,16,576,3),("in c960 8x8",32,960,8,960,3),
        ("in up c384->192 64x64",32,384,64,192,3),("in up c576->192",32,576,64,192,3),("in up c768->384",32,768,32,384,3),
        ("ch c192 32x32",100,192,32,192,3),("ch c384 16x16",100,384,16,384,3),("ch up c1152->384",100,1152,16,384,3),
        ("1x1 c192->576",100,192,32,576,1),("lin geglu",32*1024,384,0,3072,0),("lin 384",32*1024,384,0,384,0),("lin 192 k192",32*4096,192,0,192,0)]
only=os.environ.get("ONLY"); reps=int(os.environ.get("REPS","20"))
for name,B,C,H,N,k in SHAPES:
    if only and only not in name: continue
    torch.manual_seed(0)
    d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev)
    aq=ops.ActQuant(d,z,256)
    if k:
        w=torch.randn(N,C,k,k,device=dev)*0.05; x=torch.randn(B,C,H,H,device=dev); M=B*H*H; K=C*k*k
    else:
        w=torch.randn(N,C,device=dev)*0.05; x=torch.randn(B,C,device=dev); M=B; K=C
    dw=(w.flatten(1).abs().amax(1)/7.5).reshape(-1,*([1]*(w.dim()-1))); zw=torch.full_like(dw,8.)
    bias=torch.randn(N,device=dev)
    pw=ops.pack_weight(w,dw,zw,16,want_codes=True)
    nbuf=4
    if k:
        qs=[ops.act_quant_nhwc(x,aq,k//2)[0] for _ in range(nbuf)]; outs=[torch.empty(B,N,H,H,device=dev) for _ in range(nbuf)]; hw=H*H
    else:
        qs=[ops.act_quant_rows(x,aq)[0] for _ in range(nbuf)]; outs=[torch.empty(B,N,device=dev) for _ in range(nbuf)]; hw=1
    res=torch.randn_like(outs[0])
    # correctness on a slice (first 2 images / 4096 rows)
    nb = 2 if k else 4096
    if k:
        qa=ops.act_quant_nhwc(x[:nb],aq,k//2)[0]; o=torch.empty(nb,N,H,H,device=dev); r=res[:nb].contiguous()
        ops.qgemm_i8(qa,pw,d,z,o,hw,bias=bias,residual=r)
        ai=qa[:,k//2:qa.shape[1]-k//2 or None,k//2:qa.shape[2]-k//2 or None,:C].permute(0,3,1,2).double()-128.0
        wi=pw.codes.double()-8.0
        ref=F.conv2d(ai,wi,padding=k//2)*(0.03*dw.double().reshape(1,-1,1,1))+bias.double().reshape(1,-1,1,1)+r.double()
    else:
        qa=ops.act_quant_rows(x[:nb],aq)[0]; o=torch.empty(nb,N,device=dev); r=res[:nb].contiguous()
        ops.qgemm_i8(qa,pw,d,z,o,hw,bias=bias,residual=r)
        ai=qa[:,:C].double()-128.0; wi=pw.codes.reshape(N,C).double()-8.0
        ref=(ai@wi.t())*(0.03*dw.double().reshape(1,-1))+bias.double()+r.double()
    err=float((o.double()-ref).norm()/ref.norm())
    for i in range(3): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,bias=bias)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,bias=bias)
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*1e3/reps
    ress=[torch.randn_like(outs[0]) for _ in range(nbuf)]
    for i in range(3): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,bias=bias,residual=ress[i%nbuf])
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): ops.qgemm_i8(qs[i%nbuf],pw,d,z,outs[i%nbuf],hw,bias=bias,residual=ress[i%nbuf])
    e1.record(); torch.cuda.synchronize()
    us3=e0.elapsed_time(e1)*1e3/reps
    print(f"{name:24s} M={M:7d} N={N:5d} K={K:6d} rel-err {err:.1e} {'OK ' if err<2e-6 else 'BAD'} {us:8.1f} us {2*M*N*K/us/1e6:7.1f} TOP/s | +res {us3:7.1f} us", flush=True)
