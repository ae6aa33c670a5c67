import sys, os, copy
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
om=T._oracle(kind,T._qtable(qnn),cuda)
_, blk = T._run_oracle(om, args, cuda)
named=dict(qnn.named_modules())
from qdiff.quant_layer import backend
with torch.no_grad():
    for name,(a,kw,yref) in blk.items():
        if 'transformer_blocks' not in name or not any(k in name for k in ('input_blocks.4.1','input_blocks.8.1','output_blocks.2.1')): continue
        mod=named["model."+name]
        res=[]
        for label,setter in (("default",{}),("fuse_norm off",{"fuse_norm":False}),("fuse_epilogue off",{"fuse_epilogue":False}),("attention via bmm",{"fused_attention":False}),("all off",{"fuse_norm":False,"fuse_epilogue":False,"fused_attention":False})):
            prev={k:getattr(backend,k) for k in setter}
            for k,v in setter.items(): setattr(backend,k,v)
            y=mod(*a,**kw)
            for k,v in prev.items(): setattr(backend,k,v)
            res.append(f"{label}: {H.rel_l2(y,yref):.2e}")
        print(name, " | ".join(res))
