import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
from edadm import ops
from qdiff import attention as A
dev=torch.device('cuda:0')
def try_capture(name, fn):
    try:
        s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        g=torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out=fn()
        g.replay(); torch.cuda.synchronize()
        print(name,"OK", flush=True)
    except Exception as e:
        print(name,"FAILED", str(e).split("\n")[0], flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync err", e2)
BH=int(os.environ.get("BH","800")); T=1024; C=24
q=torch.randn(BH,C,T,device=dev); k=torch.randn(BH,C,T,device=dev); v=torch.randn(BH,C,T,device=dev)
d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); z0=torch.tensor([0.],device=dev); dp=torch.tensor([1/255.],device=dev)
with torch.no_grad():
    try_capture("uaq q", lambda: ops.uaq_fake_quant(q*0.5,d,z,256))
    try_capture("einsum qk", lambda: torch.einsum("bct,bcs->bts", q, k))
    try_capture("bmm qk", lambda: A.qk_scores_bct(q,k))
    w=A.qk_scores_bct(q,k)
    try_capture("softmax", lambda: torch.softmax(w.float(),dim=-1))
    p=torch.softmax(w,dim=-1)
    try_capture("uaq p", lambda: ops.uaq_fake_quant(p,dp,z0,256))
    try_capture("einsum smv", lambda: torch.einsum("bts,bcs->bct", p, v))
