"""which part of the ImageNet 32x32 transformer block carries the 1e-3 block-level difference vs the CPU oracle?"""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
import helpers as H
import test_gpu_parity_full as T
cuda=torch.device('cuda:0')
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
kind='imagenet'
qnn=T._product(kind,cuda,T._inputs(kind,4,1234))
args=T._inputs(kind,2,77)
from qdiff.quant_layer import backend
for dev in (cuda,):
    om=T._oracle(kind,T._qtable(qnn),dev)
    store={}
    hooks=[]
    for name,m in om.model.named_modules():
        if type(m).__name__ in ("CrossAttention","FeedForward","BasicTransformerBlock","LayerNorm") and "input_blocks.8.1" in name:
            def hook(mod,a,kw,out,name=name): store[name]=(tuple(x.detach() if torch.is_tensor(x) else x for x in a),{k:(v.detach() if torch.is_tensor(v) else v) for k,v in kw.items()},out.detach())
            hooks.append(m.register_forward_hook(hook,with_kwargs=True))
    with torch.no_grad(): om(*[a.to(dev) for a in args])
    named=dict(qnn.named_modules())
    print("oracle on",dev)
    with torch.no_grad():
        for name,(a,kw,yref) in store.items():
            mod=named["model."+name]
            a=tuple(t.to(cuda) if torch.is_tensor(t) else t for t in a); kw={k:(v.to(cuda) if torch.is_tensor(v) else v) for k,v in kw.items()}
            y=mod(*a,**kw)
            backend.fuse_norm=False
            y2=mod(*a,**kw)
            backend.fuse_norm=True
            print(f"  {name:55s} {type(mod).__name__:28s} rel-L2 {H.rel_l2(y.cpu(),yref.cpu()):.3e}   fuse_norm off: {H.rel_l2(y2.cpu(),yref.cpu()):.3e}")
# LN producer on the block's real input vs torch LN + quantizer
from edadm import ops
with torch.no_grad():
    tbname="input_blocks.8.1.transformer_blocks.0"
    a,kw,_=store[tbname]
    x=a[0].to(cuda)
    tb=named["model."+tbname]
    for nm,norm,lin in (("norm1->to_q",tb.norm1,tb.attn1.to_q),("norm1->to_v",tb.norm1,tb.attn1.to_v),("norm3->ff.proj",tb.norm3,tb.ff.net[0].proj)):
        aqz=lin.act_quantizer
        aq=ops.ActQuant(aqz.delta,aqz.zero_point,aqz.n_levels)
        q,_=ops.layernorm_quant_rows(x,norm.weight,norm.bias,norm.eps,aq)
        y=norm(x)
        ref=torch.clamp(torch.round(y/aqz.delta)+aqz.zero_point,0,aqz.n_levels-1).reshape(-1,y.shape[-1])
        d=(q[:,:y.shape[-1]].float()-ref).abs()
        rows=(d>0).any(1)
        print(nm,"K",y.shape[-1],"flip rate",float((d>0).float().mean()),"rows with flips",int(rows.sum()),"of",d.shape[0],"max",float(d.max()), "x mean/std", float(x.mean()), float(x.std()), "delta", float(aqz.delta))
        xr=x.reshape(-1,x.shape[-1])
        mu=xr.mean(1); sd=xr.std(1)
        print("   row |mean|/std max", float((mu.abs()/sd).max()), "median", float((mu.abs()/sd).median()))
