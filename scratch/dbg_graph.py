import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch, traceback
from edadm import ops
from oracle import qdiff_oracle as O
dev=torch.device('cuda:0')
def try_capture(name, fn):
    try:
        s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn(); fn()
        torch.cuda.current_stream().wait_stream(s)
        g=torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out=fn()
        g.replay(); torch.cuda.synchronize()
        print(name,"OK")
    except Exception as e:
        print(name,"FAILED", str(e).split("\n")[0]); torch.cuda.synchronize()
x=torch.randn(4,64,16,16,device=dev); d=torch.tensor([0.03],device=dev); z=torch.tensor([128.],device=dev)
try_capture("uaq_fwd", lambda: ops.uaq_forward(x,d,z,256))
aq=ops.ActQuant(d,z,256)
try_capture("act_quant_nhwc", lambda: ops.act_quant_nhwc(x,aq,1))
w=torch.randn(64,64,3,3,device=dev)*0.05
dw,zw,_=O.init_scale(w.cpu(),4,True)
pw=ops.pack_weight(w,dw.to(dev),zw.to(dev),16)
q,_=ops.act_quant_nhwc(x,aq,1)
out=torch.empty(4,64,16,16,device=dev)
try_capture("qgemm", lambda: ops.qgemm_i8(q,pw,d,z,out,256))
import helpers as H
from qdiff import QuantModel
g=H.load("ddim_tiny.npz")
model=H.ddim_tiny_model(); model.load_state_dict(H.state_dict(g)); model=model.to(dev)
qnn=QuantModel(model,H.WQ,H.AQ,sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); qnn.model.config.split_shortcut=True
T=torch.from_numpy
xx,tt=T(g["x"])[:4].to(dev),T(g["t"])[:4].to(dev)
with torch.no_grad():
    qnn(xx,tt); H.install_qparams(qnn,H.qtable(g)); qnn.set_quant_state(True,True)
    from qdiff.quant_layer import QuantModule
    try_capture("fp-model", lambda: model.norm_out(xx.new_zeros(4,32,16,16)))
    for n,m in list(qnn.named_modules())[:0]: pass
    try_capture("tiny unet", lambda: qnn(xx,tt))
    qnn.set_quant_state(False,False)
    try_capture("tiny unet fp", lambda: qnn(xx,tt))
