import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
import helpers as H
from qdiff import QuantModel
from qdiff.quant_layer import QuantModule, backend
T=torch.from_numpy
cuda=torch.device('cuda:0')
g=H.load("ddim_tiny.npz")
model=H.ddim_tiny_model(); model.load_state_dict(H.state_dict(g)); model=model.to(cuda)
qnn=QuantModel(model,H.WQ,H.AQ,sm_abit=8).to(cuda).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization()
qnn.model.config.split_shortcut=True
x,t=T(g["x"])[:4].to(cuda),T(g["t"])[:4].to(cuda)
outs={}
def mk(name):
    def hook(m,i,o): outs.setdefault(name,[]).append((i[0].detach().clone(), o.detach().clone()))
    return hook
for n,m in qnn.named_modules():
    if isinstance(m,QuantModule): m.register_forward_hook(mk(n))
with torch.no_grad():
    qnn(x,t)
    H.install_qparams(qnn,H.qtable(g))
    outs.clear()
    qnn.set_quant_state(True,True)
    backend.integer_path=False
    y_fake=qnn(x,t)
    fake=dict(outs); outs.clear()
    backend.integer_path=True
    y_int=qnn(x,t)
print("fake vs golden", H.rel_l2(y_fake.cpu(),T(g["y_w4a8"])), "int vs golden", H.rel_l2(y_int.cpu(),T(g["y_w4a8"])))
# per-layer: feed the SAME input (from the fake run) through the int path
backend.integer_path=True
with torch.no_grad():
    for n,m in qnn.named_modules():
        if isinstance(m,QuantModule) and n in fake:
            xin,yref=fake[n][0]
            kw={}
            y=m(xin) 
            print(f"{n:45s} {m.last_path:5s} split={m.split} in={tuple(xin.shape)} rel={H.rel_l2(y.cpu(),yref.cpu()):.2e}")
