import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
from edadm import ops
from oracle import qdiff_oracle as O
cuda=torch.device('cuda:0')
g=torch.Generator().manual_seed(0)
x=torch.softmax(torch.randn(8,64,64,generator=g)*2,-1)
gy=torch.randn(x.shape,generator=g)*1e-3
for delta in (0.00024, 0.004):
    d=torch.tensor(delta); z=torch.tensor(0.)
    xr=x.clone().requires_grad_(True); dr=d.clone().requires_grad_(True)
    O.uaq_forward(xr,dr,z,256).backward(gy)
    xc=x.to(cuda).requires_grad_(True); dc=d.to(cuda).requires_grad_(True)
    ops.uaq_fake_quant(xc,dc,z.to(cuda),256).backward(gy.to(cuda))
    print(delta, "ref gd", dr.grad.item(), "ours", dc.grad.item(), "gx equal", torch.equal(xc.grad.cpu(), xr.grad))
    # through checkpoint
    from torch.utils.checkpoint import checkpoint
    xc2=x.to(cuda).requires_grad_(True); dc2=d.to(cuda).requires_grad_(True)
    def f(inp): return ops.uaq_fake_quant(inp,dc2,z.to(cuda),256)
    checkpoint(f, xc2, use_reentrant=False).backward(gy.to(cuda))
    print("   via checkpoint", dc2.grad.item())
