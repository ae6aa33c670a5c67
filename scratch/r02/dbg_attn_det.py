import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from edadm import ops
dev=torch.device('cuda:0')
def aq(x, levels=256):
    d=(x.abs().max()*2/255).reshape(1); z=torch.tensor([128.],device=dev); return (d, z, levels)
for (B,d,T) in [(8,192,16),(8,192,64),(8,192,32),(8,64,16),(8,192,128),(8,192,256),(16,24,16),(8,192,48)]:
    g=torch.Generator().manual_seed(1)
    q=torch.randn(B,d,T,generator=g).to(dev); k=torch.randn(B,d,T,generator=g).to(dev); v=torch.randn(B,d,T,generator=g).to(dev)
    A=ops.AttnQuant(aq(q),aq(k),aq(v),(torch.tensor([1/255.],device=dev),torch.tensor([0.],device=dev),256))
    outs=[]
    for i in range(4):
        junk=torch.randn(1<<24,device=dev)*i   # perturb allocator contents
        del junk
        outs.append(ops.qattn_bct(q,k,v,A,1.0,d**-0.5).clone())
    diffs=[(outs[0]-o).abs().max().item() for o in outs[1:]]
    print((B,d,T),'run-to-run max diff',diffs, 'finite', torch.isfinite(outs[0]).all().item())
