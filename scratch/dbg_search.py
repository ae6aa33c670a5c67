import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
import helpers as H
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
T=torch.from_numpy
cuda=torch.device('cuda:0')
g=H.load("ddim_tiny.npz")
model=H.ddim_tiny_model(); model.load_state_dict(H.state_dict(g)); model=model.to(cuda)
qnn=QuantModel(model,H.WQ,H.AQ,sm_abit=8).to(cuda).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization()
qnn.model.config.split_shortcut=True
x,t=T(g["x"]).to(cuda),T(g["t"]).to(cuda)
set_weight_quantize_params(qnn,(x,t))
set_act_quantize_params(qnn,(x,t),batch_size=16)
table=H.qtable(g); named=dict(qnn.named_modules())
for name,(d,z,bits) in table.items():
    q=named[name]
    dd=q.delta.detach().cpu().reshape(-1); zz=q.zero_point.cpu().reshape(-1)
    rel=float(((dd-d.reshape(-1)).abs()/d.reshape(-1)).max())
    zeq=bool(torch.equal(zz,z.reshape(-1)))
    if rel>1e-5 or not zeq:
        nbad=int((((dd-d.reshape(-1)).abs()/d.reshape(-1))>1e-5).sum())
        print(f"{name:50s} bits={bits} n={dd.numel()} maxrel={rel:.3e} nbad={nbad} zp_equal={zeq} ours_bits={q.n_bits}")
