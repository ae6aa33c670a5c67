"""Event-timed microbenchmark of the activation producers (plain and GroupNorm-fused) at the UNet's tensor shapes."""
import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
dev=torch.device('cuda:0')
d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
def timeit(f, n=20):
    for _ in range(3): f(0)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): f(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)*1e3/n
for (B,C,H) in [(128,192,64),(100,192,32),(100,384,32),(100,384,16),(100,768,8),(128,384,32),(800,24,32)]:
    xs=[torch.randn(B,C,H,H,device=dev) for _ in range(4)]
    n=B*C*H*H
    us=timeit(lambda i: ops.act_quant_nhwc(xs[i%4],aq,1))
    a=torch.rand(B,C,device=dev)+0.5; s=torch.randn(B,C,device=dev)*0.1
    us2=timeit(lambda i: ops.norm_act_quant_nhwc(xs[i%4],a,s,True,aq,1))
    g=torch.ones(C,device=dev); b=torch.zeros(C,device=dev)
    us3=timeit(lambda i: ops.gn_fold(xs[i%4],g,b,32 if C%32==0 else 8,1e-5))
    y=torch.empty_like(xs[0])
    us4=timeit(lambda i: y.copy_(xs[i%4]))
    print(f"[{B},{C},{H},{H}] {4*n/1e6:6.1f} MB  plain {us:7.1f} us {5*n/us/1e6:5.2f} TB/s | gn+silu fused {us2:7.1f} us {5*n/us2/1e6:5.2f} TB/s | gn_fold {us3:6.1f} us {4*n/us3/1e6:5.2f} TB/s | torch copy {us4:6.1f} us {8*n/us4/1e6:5.2f} TB/s", flush=True)
