import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from edadm import ops
dev=torch.device('cuda:0')
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)*1e3/n
for (M,N,K,geglu) in [(131072,384,384,False),(32768,576,576,False),(8192,960,960,False),(131072,3072,384,True)]:
    g=torch.Generator().manual_seed(0)
    w=(torch.randn(N,K,generator=g)*0.05).to(dev)
    dw=(w.abs().amax(1)*2/15).clamp_min(1e-8); zw=torch.full((N,),8.0,device=dev)
    pw=ops.pack_weight(w, dw, zw, 16)
    qs=[torch.randint(0,256,(M,K),dtype=torch.uint8,generator=g).to(dev) for _ in range(4)]
    da=torch.tensor([0.02],device=dev); za=torch.tensor([128.],device=dev)
    cons=(torch.tensor([0.05],device=dev), torch.tensor([128.],device=dev), 256)
    i=[0]
    def run(rs):
        i[0]+=1
        return ops.qgemm_i8_codes(qs[i[0]%4], pw, da, za, cons, geglu=geglu, want_rowsum=rs)
    us1=t(lambda: run(True)); us0=t(lambda: run(False))
    line='%s codes rowsum %.1f us, no rowsum %.1f us (%.0f TOP/s)' % ((M,N,K,geglu), us1, us0, 2.0*M*N*K/us0/1e6)
    if not geglu:
        out=torch.empty(M,N,device=dev)
        usf=t(lambda: ops.qgemm_i8(qs[0], pw, da, za, out, 1))
        line += ' | fp32 rows out %.1f us' % usf
    print(line)
