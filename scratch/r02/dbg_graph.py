import sys, os, random
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch, numpy as np
import helpers as H
import test_gpu_model as TM
from qdiff_control.block_recon import block_reconstruction
from qdiff.quant_layer import backend
cuda=torch.device('cuda:0')
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
g = H.load("cfg_xattn_tiny.npz")
T=torch.from_numpy
def run(unit_name, use_graph, ckpt):
    qnn = TM._product(g, H.ldm_model("ldm_xattn_tiny.npz"), cuda, TM._set_split_ldm)
    cali = tuple(T(g[k]).to(cuda) for k in ("x", "t", "index", "cond", "uncond"))
    with torch.no_grad(): qnn(cali[0][:4], cali[1][:4], cali[3][:4])
    H.install_qparams(qnn, H.qtable(g))
    kw = dict(TM.RECON_KW); kw.update(batch_size=4, iters=8)
    random.seed(56); torch.manual_seed(56)
    unit = qnn.model.input_blocks[1][0] if unit_name=='res' else qnn.model.input_blocks[1][1].transformer_blocks[0]
    if unit_name=='tb': unit.checkpoint = ckpt
    timing={"warmup":0,"grad_norms":[]}
    backend.recon_cuda_graph = use_graph
    losses = block_reconstruction(qnn, unit, cali_data=cali, return_losses=True, timing=timing, **kw)
    backend.recon_cuda_graph = True
    return timing["cuda_graph"], losses.cpu().numpy(), torch.stack(timing["grad_norms"]).cpu().numpy().T
for unit,ck in (('tb',True),):
    for ug in (True,True,False,False):
        print(unit,'ckpt',ck,'graph',ug, run(unit,ug,ck))
