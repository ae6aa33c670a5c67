import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200'), os.path.join(ROOT,'tests')]
import torch
import bench
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
from qdiff.quant_layer import QuantModule
dev=torch.device('cuda:0')
kind,batch,shape,ctx,_=bench.WORKLOADS["church"]
batch=int(os.environ.get("B","100"))
fp=bench.build_fp_unet(kind).to(dev)
qnn=QuantModel(fp,bench.WQ,bench.AQ,sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model,kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape,ctx,64,1234)]
set_weight_quantize_params(qnn,cali); set_act_quantize_params(qnn,cali,batch_size=32)
qnn.set_quant_state(True,True)
xin=[c.to(dev) for c in bench.synth_inputs(shape,ctx,batch,5)]
# record per-top-level-module inputs
rec={}
def mk(n):
    def pre(m,args,kwargs): rec[n]=(args,kwargs)
    return pre
tops=[("time_embed",qnn.model.time_embed)]+[(f"input_blocks.{i}",b) for i,b in enumerate(qnn.model.input_blocks)]+[("middle",qnn.model.middle_block)]+[(f"output_blocks.{i}",b) for i,b in enumerate(qnn.model.output_blocks)]+[("out",qnn.model.out)]
hs=[m.register_forward_pre_hook(mk(n),with_kwargs=True) for n,m in tops]
with torch.no_grad(): qnn(*xin)
for h in hs: h.remove()
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
        print(name,"OK", flush=True)
        return True
    except Exception as e:
        print(name,"FAILED", str(e).split("\n")[0], flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync err", e2)
        return False
with torch.no_grad():
    for n,m in tops:
        a,k=rec[n]
        ok=try_capture(n, lambda: m(*a,**k))
        if not ok:
            # bisect into children QuantModules
            sub={}
            hs=[mm.register_forward_pre_hook((lambda nn_: (lambda mod,args,kwargs: sub.__setitem__(nn_,(args,kwargs))))(nn_),with_kwargs=True) for nn_,mm in m.named_modules() if isinstance(mm,QuantModule)]
            m(*a,**k)
            for h in hs: h.remove()
            for nn_,mm in m.named_modules():
                if isinstance(mm,QuantModule):
                    aa,kk=sub[nn_]
                    print("   ", nn_, tuple(aa[0].shape), mm.fwd_kwargs, end=" ")
                    try_capture("", lambda: mm(*aa,**kk))
            break
