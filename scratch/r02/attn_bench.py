import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from edadm import ops
dev=torch.device('cuda:0')
def aq(x, levels=256):
    d=(x.abs().max()*2/255).reshape(1); z=torch.tensor([128.],device=dev); return (d, z, levels)
for (BH,heads,d,T) in [(128,1,384,1024),(128,1,576,256),(128,1,960,64),(800,8,24,1024),(800,8,48,256),(800,8,96,64)]:
    g=torch.Generator().manual_seed(1)
    q=torch.randn(BH,T,d,generator=g).to(dev); k=torch.randn(BH,T,d,generator=g).to(dev); v=torch.randn(BH,T,d,generator=g).to(dev)
    A=ops.AttnQuant(aq(q),aq(k),aq(v),(torch.tensor([1/255.],device=dev),torch.tensor([0.],device=dev),256))
    f=lambda: ops.qattn_bnd(q,k,v,heads,A,d**-0.5)
    o=f(); 
    for _ in range(2): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize(); us=e0.elapsed_time(e1)*100
    print((BH,heads,d,T), '%.1f us incl. producers' % us, 'checksum %.6f' % o.double().sum().item())
