"""Kernel-time breakdown of eager block-reconstruction iterations (torch.profiler, CUDA activities)."""
import sys, os, collections, re
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import torch, bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.block_recon import block_reconstruction
from qdiff.quant_layer import backend
from qdiff.quant_block import BaseQuantBlock
wl=os.environ.get("WL","church")
kind, batch, shape, ctx, _ = bench.WORKLOADS[wl]
dev=torch.device("cuda:0")
fp=bench.build_fp_unet(kind).to(dev)
qnn=QuantModel(fp, bench.WQ, bench.AQ, sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape, ctx, 64, seed=1234)]
set_weight_quantize_params(qnn, cali); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True)
units=[m for m in qnn.model.modules() if isinstance(m, BaseQuantBlock) and type(m).__name__ in ("QuantResBlock","QuantResnetBlock")]
unit=units[len(units)//4]
backend.recon_cuda_graph=False
cali_r=[c.to(dev) for c in bench.synth_inputs(shape, ctx, 64, seed=4321)]
kw=dict(cali_data=cali_r, iters=13, batch_size=32, weight=0.01, asym=True, b_range=(20,2), warmup=0.2, act_quant=True, opt_mode='mse',
        lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=0.5, keep_gpu=True, recon_w=True, recon_a=True, add_loss=0.8)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    block_reconstruction(qnn, unit, **kw)
agg=collections.defaultdict(lambda:[0,0.0])
for e in prof.events():
    if e.device_type.name=="CUDA":
        n=re.sub(r'<.*','',e.name); n=re.sub(r'\(.*','',n)[:70]
        agg[n][0]+=1; agg[n][1]+=e.device_time if hasattr(e,'device_time') else e.cuda_time
tot=sum(v[1] for v in agg.values())
print(f"{wl}: total CUDA time {tot/1e3:.1f} ms over the whole call (cache build + 13 iterations)")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{v[1]/1e3:9.2f} ms {v[0]:6d}  {100*v[1]/tot:5.1f}%  {k}")
