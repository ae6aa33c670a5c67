import sys, torch
sys.path[:0]=['.', 'eda-dm_b200']
from edadm import ops
dev=torch.device('cuda:0')
def t(f):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)*100
out=[]
for shape in ((128,192,64,64),(128,384,32,32),(128,576,16,16),(128,960,8,8)):
    x=torch.randn(*shape,device=dev); d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
    a_,s_=ops.gn_fold(x,torch.randn(shape[1],device=dev),torch.randn(shape[1],device=dev),32,1e-5)
    us1=t(lambda: ops.norm_act_quant_nhwc(x,a_,s_,1,aq,1)); us0=t(lambda: ops.act_quant_nhwc(x, aq, 1))
    out.append('%s silu %.1f us %.0f GB/s | plain %.1f us %.0f GB/s' % (shape, us1, x.numel()*5/us1/1e3, us0, x.numel()*5/us0/1e3))
print('\n'.join(out))
