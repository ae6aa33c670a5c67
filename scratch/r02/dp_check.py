"""2-GPU check: overlapped per-parameter all-reduce == single bucket all-reduce (bit-identical alphas / step sizes)"""
import os, sys, random
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch, torch.distributed as dist
import helpers as H
import test_gpu_model as TM
from qdiff.block_recon import block_reconstruction
from qdiff.quant_layer import backend
rank=int(os.environ["RANK"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32=False; torch.backends.cuda.matmul.allow_tf32=False
g=H.load("ddim_tiny.npz"); T=torch.from_numpy
res=[]
for overlap in (True, True, False, False):
    backend.recon_overlap_allreduce=overlap
    qnn=TM._product(g,H.ddim_tiny_model(),dev,TM._set_split_ddim)
    x,t=T(g["x"]).repeat(2,1,1,1).to(dev),T(g["t"]).repeat(2).to(dev)
    with torch.no_grad(): qnn(x[:4],t[:4])
    H.install_qparams(qnn,H.qtable(g))
    random.seed(77); torch.manual_seed(77)
    blk=qnn.model.down[0].block[0]
    kw=dict(TM.RECON_KW); kw.update(iters=8,batch_size=4)
    losses=block_reconstruction(qnn,blk,cali_data=(x,t),return_losses=True,**kw)
    res.append((losses.cpu(), blk.conv1.weight_quantizer.alpha.detach().cpu().clone(), blk.conv2.act_quantizer.delta.detach().cpu().clone()))
same = (torch.equal(res[0][1],res[1][1]), torch.equal(res[2][1],res[3][1]), torch.equal(res[0][1],res[2][1]), float((res[0][1]-res[2][1]).abs().max()))
# replicas must hold identical parameters
a=res[0][1].to(dev); b=a.clone(); dist.broadcast(b,0)
print(f"rank {rank}: overlap == bucket: {same}; losses {res[0][0].tolist()[:4]}; replicas identical: {bool(torch.equal(a,b))}", flush=True)
dist.destroy_process_group()
