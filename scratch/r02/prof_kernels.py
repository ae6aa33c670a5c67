"""Launch one kernel family a few times at BASELINE-like sizes (driver for `ncu --set full -k regex:...`).  usage: prof_kernels.py <which>"""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
dev=torch.device('cuda:0'); which=sys.argv[1]
g=torch.Generator().manual_seed(0)
R=lambda *s: torch.randn(*s, generator=g).to(dev)
d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
def gemm_setup(B,C,H,N,k):
    w=R(N,C,k,k)*0.05 if k else R(N,C)*0.05
    dw=(w.flatten(1).abs().amax(1)/7.5).reshape(-1,*([1]*(w.dim()-1))); zw=torch.full_like(dw,8.)
    return ops.pack_weight(w,dw,zw,16,w4=False)
for rep in range(4):
    if which=='uaq':                      # fake-quant fwd + bwd (QDrop draws), 32x224x64x64 (bedroom L0 activation)
        x=R(32,224,64,64).requires_grad_(True); dl=torch.nn.Parameter(torch.tensor(0.03,device=dev))
        y=ops.uaq_fake_quant(x,dl,z,256,None,0.5,0,0,keep_rand=torch.rand_like(x)); y.backward(torch.ones_like(y))
    elif which=='adaround':               # 24 M weights (largest bedroom unit)
        w=R(24*1024*1024//64,64)*0.05; dw=(w.abs().amax(1)/7.5).reshape(-1,1); zw=torch.full_like(dw,8.)
        al=ops.adaround_init_alpha(w,dw).requires_grad_(True)
        y=ops.adaround_fake_quant(w,al,dw,zw,16,True); y.backward(torch.ones_like(y))
    elif which=='lp_loss':
        a=R(32,224,64,64).requires_grad_(True); b=R(32,224,64,64); l=ops.lp_loss(a,b,2.0); l.backward()
    elif which=='gn_fold':
        x=R(128,192,64,64); gm=R(192); bt=R(192); ops.gn_fold(x,gm,bt,32,1e-5)
    elif which=='actq':                   # fused GroupNorm + SiLU (exact) + quantize producer, ImageNet 64x64 level
        x=R(128,192,64,64); a_,s_=ops.gn_fold(x,R(192),R(192),32,1e-5); ops.norm_act_quant_nhwc(x,a_,s_,1,aq,1)
    elif which=='layernorm':
        x=R(128*1024,384); ln=torch.nn.LayerNorm(384).to(dev); ops.layernorm_quant_rows_multi(x,ln.weight,ln.bias,ln.eps,[aq,aq,aq],[False]*3)
    elif which=='rows':
        ops.act_quant_rows(R(128*1024,384),aq)
    elif which=='conv_small':
        x=R(128,192,64,64); a_,s_=ops.gn_fold(x,R(192),R(192),32,1e-5); ops.conv3x3_small_n(x,R(3,192,3,3)*0.05,R(3),affine=(a_,s_,1))
    elif which=='search':
        x=R(32,384,32,32).reshape(1,-1); dl=torch.linspace(0.01,0.05,100,device=dev).reshape(1,-1); zz=torch.full_like(dl,128.)
        ops.mse_search_scores(x,dl,zz,256)
    elif which in ('gemm_c384','gemm_c192','gemm_up'):
        B,C,H,N={'gemm_c384':(128,384,32,384),'gemm_c192':(128,192,64,192),'gemm_up':(128,384,64,192)}[which]
        pw=gemm_setup(B,C,H,N,3); q,_=ops.act_quant_nhwc(R(B,C,H,H),aq,1); out=torch.empty(B,N,H,H,device=dev)
        ops.qgemm_i8(q,pw,d,z,out,H*H)
    elif which=='gemm_geglu':
        pw=gemm_setup(0,384,0,3072,0); q,_=ops.act_quant_rows(R(128*1024,384),aq)
        ops.qgemm_i8_codes(q,pw,d,z,(d,z,256),geglu=True)
    elif which=='gemm_lin':
        pw=gemm_setup(0,1536,0,384,0); q,_=ops.act_quant_rows(R(128*1024,1536),aq); out=torch.empty(128*1024,384,device=dev); res=R(128*1024,384)
        ops.qgemm_i8(q,pw,d,z,out,1,residual=res)
    elif which in ('attn_in','attn_church'):
        BH,heads,dd,T={'attn_in':(128,1,384,1024),'attn_church':(800,8,24,1024)}[which]
        q_=R(BH,T,dd); k_=R(BH,T,dd); v_=R(BH,T,dd)
        aqn=ops.AttnQuant((d,z,256),(d,z,256),(d,z,256),(torch.tensor([1/255.],device=dev),torch.tensor([0.],device=dev),256))
        ops.qattn_bnd(q_,k_,v_,heads,aqn,dd**-0.5)
    elif which=='bf16x3_conv':            # reconstruction-loop conv on the bf16 x 3 kernel: ImageNet 16x16 ResBlock conv 576->576, batch 32
        x=R(32,576,16,16); w=R(576,576,3,3)*0.03; ops.conv_bf16x3(x,w,R(576))
    elif which=='bf16x3_wgrad':           # convolution wgrad of the same ResBlock conv (pixels = reduction dimension, split-K)
        x=R(32,576,16,16); gy=R(32,576,16,16); ops.conv_wgrad_bf16x3(gy,x,3)
    elif which=='bf16x3_lin':             # transformer-block linear of the reconstruction loop: 32 x 1024 tokens, 384 -> 3072
        x=R(32*1024,384); w=R(3072,384)*0.05; ops.linear_bf16x3(x,w,R(3072))
    torch.cuda.synchronize()
