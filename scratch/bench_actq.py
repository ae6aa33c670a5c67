import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
dev=torch.device('cuda:0')
d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
for (B,C,H) in [(128,192,64),(100,192,32),(128,384,32),(800,24,32)]:
    xs=[torch.randn(B,C,H,H,device=dev) for _ in range(3)]
    for i in range(3): ops.act_quant_nhwc(xs[i%3],aq,1)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): ops.act_quant_nhwc(xs[i%3],aq,1)
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*100
    n=B*C*H*H
    print(f"nhwc [{B},{C},{H},{H}] {us:8.1f} us  {5*n/us/1e3:7.1f} GB/s (incl. halo kernel + alloc)")
    y=torch.empty_like(xs[0])
    e0.record()
    for i in range(10): y.copy_(xs[i%3])
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*100
    print(f"   torch copy  {us:8.1f} us {8*n/us/1e3:7.1f} GB/s")
x=torch.randn(131072,384,device=dev)
for i in range(3): ops.act_quant_rows(x,aq)
torch.cuda.synchronize(); e0.record()
for i in range(10): ops.act_quant_rows(x,aq)
e1.record(); torch.cuda.synchronize()
us=e0.elapsed_time(e1)*100
print(f"rows [131072,384] {us:8.1f} us {5*x.numel()/us/1e3:7.1f} GB/s")
