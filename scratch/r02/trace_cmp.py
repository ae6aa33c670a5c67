import sys, os, random, numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import helpers as H
from test_gpu_model import _product, _set_split_ddim, RECON_KW, T
from qdiff.block_recon import block_reconstruction
from qdiff.quant_layer import backend
cuda=torch.device('cuda:0')
g = H.load("ddim_tiny.npz")
ref = g["recon_block_loss"]
def run(tag, bf16x3, tf32, iters=4):
    qnn = _product(g, H.ddim_tiny_model(), cuda, _set_split_ddim)
    x, t = T(g["x"]).to(cuda), T(g["t"]).to(cuda)
    with torch.no_grad(): qnn(x[:4], t[:4])
    H.install_qparams(qnn, H.qtable(g))
    random.seed(77); torch.manual_seed(77)
    backend.calib_gemm_bf16x3 = bf16x3; backend.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32 if tf32 is not None else True
    torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    losses = block_reconstruction(qnn, qnn.model.down[0].block[0], cali_data=(x, t), return_losses=True, **dict(RECON_KW, iters=iters)).cpu().numpy()
    print(tag, losses, "rel dev", np.abs(losses/ref[:iters]-1))
run("bf16x3 fwd/dgrad + ambient(TF32) wgrad", True, False)
torch.backends.cudnn.allow_tf32=False
run("bf16x3 + fp32 wgrad", True, None) if False else None
run("library fp32 fwd + TF32 bwd (round-1 state)", False, False)
run("library TF32 everywhere (PyTorch default = the reference on a GPU)", False, True)
