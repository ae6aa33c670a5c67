import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
import helpers as H
from qdiff import QuantModel
from qdiff.quant_layer import QuantModule, backend
from oracle.model_oracle import OracleQuantUNet, OQuantLayer
T=torch.from_numpy
cuda=torch.device('cuda:0')
g=H.load("ddim_tiny.npz")
model=H.ddim_tiny_model(); model.load_state_dict(H.state_dict(g)); model=model.to(cuda)
qnn=QuantModel(model,H.WQ,H.AQ,sm_abit=8).to(cuda).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization()
qnn.model.config.split_shortcut=True
x,t=T(g["x"])[:4],T(g["t"])[:4]
outs={}
def mk(store,name):
    def hook(m,i,o): store[name]=(i[0].detach().cpu().clone(), o.detach().cpu().clone())
    return hook
for n,m in qnn.named_modules():
    if isinstance(m,QuantModule): m.register_forward_hook(mk(outs,n))
m2=H.ddim_tiny_model(); m2.load_state_dict(H.state_dict(g))
om=OracleQuantUNet(m2,H.WQ,H.AQ,8); om.set_first_last_layer_to_8bit(); om.disable_network_output_quantization(); m2.config.split_shortcut=True
oo={}
for n,l in om.layers: l.register_forward_hook(mk(oo,n))
with torch.no_grad():
    qnn(x.to(cuda),t.to(cuda)); om(x,t)
    H.install_qparams(qnn,H.qtable(g)); om.load_qparams(H.qtable_for_oracle(g,om))
    qnn.set_quant_state(True,True); om.set_quant_state(True,True)
    backend.integer_path=False
    y=qnn(x.to(cuda),t.to(cuda)); yo=om(x,t)
print("oracle vs golden", H.rel_l2(yo,T(g["y_w4a8"])), "gpu fake vs golden", H.rel_l2(y.cpu(),T(g["y_w4a8"])))
for n in outs:
    i1,o1=outs[n]; i2,o2=oo[n]
    print(f"{n:45s} in rel={H.rel_l2(i1,i2):.2e} maxabs={float((i1-i2).abs().max()):.2e} out rel={H.rel_l2(o1,o2):.2e}")
