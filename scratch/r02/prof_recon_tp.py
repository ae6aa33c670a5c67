"""kernel time table of eager reconstruction iterations of an ImageNet unit (torch.profiler)"""
import sys, os, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import backend
from qdiff_control.block_recon import block_reconstruction
dev=torch.device('cuda:0'); which=sys.argv[1]
kind,batch,shape,ctx,_=bench.WORKLOADS['imagenet']
fp = bench.build_fp_unet(kind).to(dev)
qnn = QuantModel(fp, bench.WQ, bench.AQ, sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape,ctx,64,seed=1234)]
set_weight_quantize_params(qnn, cali); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True)
unit = bench.recon_units(qnn, kind)[which][0]
g = torch.Generator().manual_seed(99)
x,t=[c.to(dev) for c in bench.synth_inputs(shape,None,64,seed=4321)]
cond=torch.randn(64,*ctx,generator=g).to(dev); uncond=torch.randn(64,*ctx,generator=g).to(dev)
backend.recon_cuda_graph=False
kw = dict(cali_data=(x,t,torch.zeros(64,dtype=torch.long,device=dev),cond,uncond), iters=10, batch_size=32, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
          act_quant=True, opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=0.5, keep_gpu=True, recon_w=True, recon_a=True, add_loss=0.8)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    block_reconstruction(qnn, unit, **kw)
    torch.cuda.synchronize()
rows=sorted(prof.key_averages(), key=lambda e:-e.device_time_total)
tot=sum(e.device_time_total for e in rows)
print("total device us", tot)
for e in rows[:28]:
    print("%9.1f us %5.1f%% x%-4d %s" % (e.device_time_total, 100*e.device_time_total/tot, e.count, e.key[:110]))
