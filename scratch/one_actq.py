import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch
from edadm import ops
dev=torch.device('cuda:0')
d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev); aq=ops.ActQuant(d,z,256)
x=torch.randn(128,192,64,64,device=dev)
for i in range(3): ops.act_quant_nhwc(x,aq,1)
torch.cuda.synchronize()
