import sys, os, random
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch, numpy as np
import helpers as H
import test_gpu_model as TM
from qdiff_control.block_recon import block_reconstruction
from qdiff.quant_layer import backend
cuda=torch.device('cuda:0'); T=torch.from_numpy
g=H.load("cfg_xattn_tiny.npz")
for graph in (True, False):
    backend.recon_cuda_graph=graph
    qnn=TM._product(g,H.ldm_model("ldm_xattn_tiny.npz"),cuda,TM._set_split_ldm)
    cali=tuple(T(g[k]).to(cuda) for k in ("x","t","index","cond","uncond"))
    with torch.no_grad(): qnn(cali[0][:4],cali[1][:4],cali[3][:4])
    H.install_qparams(qnn,H.qtable(g))
    kw=dict(TM.RECON_KW); kw.update(batch_size=4)
    random.seed(55); torch.manual_seed(55)
    res=qnn.model.input_blocks[1][0]
    l=block_reconstruction(qnn,res,cali_data=cali,return_losses=True,**kw)
    print("graph",graph,"res ours",l.cpu().numpy(),"ref",g["recon_res_loss"])
    random.seed(56); torch.manual_seed(56)
    tb=qnn.model.input_blocks[1][1].transformer_blocks[0]
    l=block_reconstruction(qnn,tb,cali_data=cali,return_losses=True,**kw)
    print("graph",graph,"tb ours",l.cpu().numpy(),"ref",g["recon_tb_loss"])
    d=[float(tb.attn1.act_quantizer_q.delta),float(tb.attn1.act_quantizer_w.delta),float(tb.attn2.act_quantizer_k.delta),float(tb.attn2.act_quantizer_v.delta)]
    print(d, g["recon_tb_delta"])
