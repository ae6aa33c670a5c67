import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
dev=torch.device('cuda:0')
def aq():
    mk=lambda d,z:(torch.tensor([d],device=dev),torch.tensor([z],device=dev),256)
    return ops.AttnQuant(mk(0.03,128.),mk(0.03,128.),mk(0.03,128.),mk(1/255.,0.))
import os
SH=[(800,24,1024)] if os.environ.get('ONE') else [(800,24,1024),(800,48,256),(128,384,1024),(128,576,256)]
for (BH,d,T) in SH:
    q,k,v=(torch.randn(BH,d,T,device=dev) for _ in range(3))
    A=aq()
    from edadm.ops import _codes_token_major_from_bct,_codes_rows,_f32c
    qc,rq=_codes_token_major_from_bct(q,A.q,1.0); kc,rk=_codes_token_major_from_bct(k,A.k,1.0); vc,rv=_codes_rows(v.reshape(BH*d,T),A.v)
    out=torch.empty(BH,d,T,device=dev)
    f=lambda: ops.qattn(qc,kc,vc.reshape(BH,d,-1),rq,rk,rv.reshape(BH,d),1,d,T,A,0.2,out,(d*T,0,1,T))
    for i in range(3): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us=e0.elapsed_time(e1)*100
    print(f"qattn BH={BH} d={d} T={T}: {us:8.1f} us  {4*BH*T*T*d/us/1e6:7.1f} TOP/s(2 GEMMs)  {BH*T*T/us/1e3:6.2f} Gscore/s")
