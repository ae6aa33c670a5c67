import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
from edadm import ops
from edadm.native import lib
import torch.nn.functional as F
dev=torch.device('cuda:0')
B,C,N,H,W,R=[int(v) for v in sys.argv[1:7]]; splits=int(sys.argv[7])
g=torch.Generator().manual_seed(0)
x=torch.randn(B,C,H,W,generator=g).to(dev); gy=torch.randn(B,N,H,W,generator=g).to(dev)
gh,gl,_,_=ops.split_bf16(gy.reshape(B*N,H*W)); xh=torch.empty(R,B,C,H,W,dtype=torch.bfloat16,device=dev); xl=torch.empty_like(xh); lib.split_shift_bf16(x.data_ptr(),xh.data_ptr(),xl.data_ptr(),B*C*H,W,R,(R-1)//2,torch.cuda.current_stream().cuda_stream)
out=torch.zeros(R*R,N,C,device=dev)
lib.conv_wgrad_bf16x3(gh.data_ptr(),gl.data_ptr(),xh.data_ptr(),xl.data_ptr(),B,N,C,H,W,R,R,(R-1)//2,out.data_ptr(),splits,torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
dw=out.permute(1,2,0).reshape(N,C,R,R)
xd=x.double(); wd=torch.zeros(N,C,R,R,dtype=torch.float64,device=dev,requires_grad=True)
F.conv2d(xd,wd,padding=(R-1)//2).backward(gy.double())
err=(dw.double()-wd.grad).norm()/wd.grad.norm()
print((B,C,N,H,W,R,splits),'rel err',err.item())
