import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
from qdiff.quant_layer import UniformAffineQuantizer
import helpers as H
torch.manual_seed(0)
w=torch.randn(24,16,3,3)*0.08
q=UniformAffineQuantizer(**H.WQ); d,z=q.init_quantization_scale_1(w,True)
q2=UniformAffineQuantizer(**H.WQ); d2,z2=q2.init_quantization_scale_1(w.cuda(),True)
print("cpu vs cuda:", float((d-d2.cpu()).abs().max()/d.abs().max()), int((z!=z2.cpu()).sum()))
# candidate scores for channel 0 on both devices
for dev in ("cpu","cuda"):
    ww=w.to(dev); y=torch.flatten(ww,1)
    qq=UniformAffineQuantizer(**H.WQ); qq.one_side_dist='no'
    xr=torch.max(y.amin(1).abs(), y.amax(1))
    sc=[]
    for i in range(1,101):
        th=xr/100*i
        sc.append(qq._score_candidates(y,-th,th,True)[0].item())
    import numpy as np
    sc=np.array(sc); print(dev, sc.argmin(), sc[sc.argsort()[:4]], sc.argsort()[:4])
