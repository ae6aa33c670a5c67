import sys, os, logging, traceback
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
logging.basicConfig(level=logging.WARNING)
import torch, bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.block_recon import block_reconstruction
from qdiff.quant_block import BaseQuantBlock
import qdiff._recon_engine as E
dev=torch.device('cuda:0')
from qdiff.quant_layer import backend
backend.allow_tf32=bool(int(os.environ.get('TF32','0')))
PROB=float(os.environ.get('PROB','0.5'))
bench.AQ['prob']=PROB
kind,batch,shape,ctx,_=bench.WORKLOADS[os.environ.get("WL","church")]
fp=bench.build_fp_unet(kind).to(dev)
qnn=QuantModel(fp,bench.WQ,bench.AQ,sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model,kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape,ctx,64,1234)]
set_weight_quantize_params(qnn,cali); set_act_quantize_params(qnn,cali,batch_size=32,all_attention=True)
units=[m for m in qnn.model.modules() if isinstance(m,BaseQuantBlock) and type(m).__name__ in ("QuantResBlock","QuantResnetBlock")]
unit=units[len(units)//4]
print(type(unit).__name__, [ (n,tuple(m.weight.shape)) for n,m in unit.named_modules() if hasattr(m,'weight') and hasattr(m,'split')])
orig=torch.cuda.graph.__exit__
timing={"warmup":3}
kw=dict(cali_data=cali, iters=30, batch_size=32, weight=0.01, asym=True, b_range=(20,2), warmup=0.2, act_quant=True, opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=PROB, keep_gpu=True, recon_w=True, recon_a=True, add_loss=0.8, timing=timing)
import qdiff._recon_engine as E
_old=E.logger.warning
def w(msg,*a):
    print("WARNING:", msg % a); traceback.print_exc()
E.logger.warning=w
try:
    block_reconstruction(qnn, unit, **kw)
    print(timing.get('ms_per_iter'), timing.get('cuda_graph'))
except Exception as e:
    traceback.print_exc()
