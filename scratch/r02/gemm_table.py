"""per-shape table of the int8 GEMM launches of one eager ImageNet step (CUDA events around every launch)"""
import sys, os, collections, torch
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0]=[ROOT, os.path.join(ROOT,'eda-dm_b200')]
import bench
from edadm import ops
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
dev=torch.device('cuda:0')
wl = sys.argv[1] if len(sys.argv) > 1 else 'imagenet'
kind,batch,shape,ctx,_=bench.WORKLOADS[wl]
fp = bench.build_fp_unet(kind).to(dev)
qnn = QuantModel(fp, bench.wq_params(kind), bench.AQ, sm_abit=8).to(dev).eval()
qnn.set_first_last_layer_to_8bit(); qnn.disable_network_output_quantization(); bench.set_split(qnn.model, kind)
cali=[c.to(dev) for c in bench.synth_inputs(shape,ctx,64,seed=1234)]
set_weight_quantize_params(qnn, cali); set_act_quantize_params(qnn, cali, batch_size=32, all_attention=True)
qnn.set_quant_state(True, True)
args=[c.to(dev) for c in bench.synth_inputs(shape,ctx,batch,seed=7)]
rec=[]
def wrap(name, fn, describe):
    def w(*a, **k):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); r=fn(*a, **k); e1.record(); rec.append((describe(*a, **k), e0, e1)); return r
    setattr(ops, name, w)
def d_gemm(q, pw, da, za, out, out_hw, bias=None, rowsum=None, a_c_offset=0, accumulate=False, silu=False, filter_rs=None, residual=None, bias_img=None):
    R,S = filter_rs if filter_rs is not None else (pw.R, pw.S)
    M = q.shape[0] if q.dim()==2 else q.shape[0]*(q.shape[1]-R+1)*(q.shape[2]-S+1)
    return ('conv' if out_hw>1 else 'lin', M, pw.N, pw.C*R*S if filter_rs is None else pw.C*pw.R*pw.S, R, 'res' if residual is not None else ('acc' if accumulate else ('bimg' if bias_img is not None else '-')))
def d_codes(q, pw, da, za, consumer, bias=None, rowsum=None, geglu=False, want_rowsum=False):
    return ('codes-geglu' if geglu else 'codes', q.shape[0], pw.N, pw.C, 1, '-')
def d_post(q, pw, da, za, out, bias, residual, post, post_rows):
    return ('lin', q.shape[0], pw.N, pw.C, 1, 'res+post')
wrap('qgemm_i8', ops.qgemm_i8, d_gemm); wrap('qgemm_i8_codes', ops.qgemm_i8_codes, d_codes); wrap('qgemm_i8_rows_post', ops.qgemm_i8_rows_post, d_post)
with torch.no_grad():
    for _ in range(2): qnn(*args)
    rec.clear()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); qnn(*args); e1.record()
torch.cuda.synchronize()
agg=collections.OrderedDict()
for d,a,b in rec:
    t=a.elapsed_time(b)*1e3
    v=agg.setdefault(d,[0,0.0]); v[0]+=1; v[1]+=t
tot=sum(v[1] for v in agg.values())
print('eager step %.2f ms, GEMM launches %d, GEMM time %.2f ms' % (e0.elapsed_time(e1), len(rec), tot/1e3))
print('%-12s %8s %6s %6s %2s %-8s %4s %9s %8s %7s %6s' % ('kind','M','N','K','R','epi','n','total us','us/launch','TOP/s','share'))
for d,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    ops_=2.0*d[1]*d[2]*d[3]
    print('%-12s %8d %6d %6d %2d %-8s %4d %9.1f %8.1f %7.0f %5.1f%%' % (d[0],d[1],d[2],d[3],d[4],d[5],v[0],v[1],v[1]/v[0],ops_/(v[1]/v[0]*1e-6)/1e12,100*v[1]/tot))
