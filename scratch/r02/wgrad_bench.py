import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from edadm import ops
from edadm.native import lib
dev=torch.device('cuda:0')
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)*1e3/n
for (B,C,N,H,W,R) in [(32,576,576,16,16,3),(32,384,384,32,32,3),(32,192,192,64,64,3),(32,128,128,32,32,3),(32,256,256,16,16,3)]:
    g=torch.Generator().manual_seed(0)
    x=torch.randn(B,C,H,W,generator=g).to(dev); gy=torch.randn(B,N,H,W,generator=g).to(dev); w=torch.randn(N,C,R,R,generator=g).to(dev)
    pad=(R-1)//2
    full=t(lambda: ops.conv_wgrad_bf16x3(gy,x,R))
    HW=H*W
    def prep():
        gh,gl,_,_=ops.split_bf16(gy.reshape(B*N,HW))
        xh=torch.empty(R,B,C,H,W,dtype=torch.bfloat16,device=dev); xl=torch.empty_like(xh)
        lib.split_shift_bf16(x.data_ptr(),xh.data_ptr(),xl.data_ptr(),B*C*H,W,R,pad,torch.cuda.current_stream().cuda_stream)
        return gh,gl,xh,xl
    tp=t(prep)
    torch.backends.cudnn.allow_tf32=True
    tf=t(lambda: torch.ops.aten.convolution_backward(gy,x,w,None,[1,1],[pad,pad],[1,1],False,[0,0],1,[False,True,False]))
    torch.backends.cudnn.allow_tf32=False
    f32=t(lambda: torch.ops.aten.convolution_backward(gy,x,w,None,[1,1],[pad,pad],[1,1],False,[0,0],1,[False,True,False]))
    fl=2.0*B*HW*N*C*R*R
    print((B,C,N,H,W,R),'bf16x3 wgrad total %.1f us (operand prep %.1f) = %.0f TFLOP/s x3 | cuDNN TF32 %.1f us | cuDNN fp32 %.1f us' % (full,tp,3*fl/(full-tp)/1e6,tf,f32))
