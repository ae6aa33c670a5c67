"""profiles/ncu_*_r02_summary.txt from gpurun_out/r02/ncu/*.ncu-rep (+ achieved GB/s or TOP/s against the measured peaks)"""
import sys, os, re
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "scratch"))
sys.argv=[sys.argv[0], "r02"]
import summarize_ncu as S
HBM=6546.2
# name -> (note, algorithmic bytes per launch or None, algorithmic ops per launch or None)
E=32*224*64*64
W=24*1024*1024
IN=128*192*64*64
CASES={
 "uaq_fwd": ("UniformAffineQuantizer forward with QDrop draws (quant_layer.py:267-274), 32x224x64x64 fp32: read x + rand, write y", E*12, None),
 "uaq_bwd": ("STE backward + step-size gradient, same tensor: read gy, x, rand, write gx", E*16, None),
 "adaround_fwd": ("AdaRound soft forward (adaptive_rounding.py:49-59), 24 M weights: read w, alpha, write w~", W*12, None),
 "adaround_bwd": ("AdaRound backward: read g, w, alpha, write g_alpha", W*16, None),
 "lp_loss_fwd": ("lp_loss forward (quant_layer.py:26-33), 32x224x64x64: read pred, tgt", E*8, None),
 "lp_loss_bwd": ("lp_loss backward: read pred, tgt, write g_pred", E*12, None),
 "gn_fold": ("GroupNorm statistics -> per-(sample, channel) affine, 128x192x64x64: one read", IN*4, None),
 "actq_exact_silu": ("fused GroupNorm + exact SiLU + quantize producer (NCHW fp32 -> NHWC u8 + halo), 128x192x64x64: read 4 B, write 1 B", IN*5, None),
 "layernorm_multi": ("LayerNorm + three quantizers in one pass (norm1 -> to_q / to_k / to_v), 131072 x 384: read 4 B, write 3 x 1 B", 131072*384*7, None),
 "conv3x3_small_n": ("output head: GroupNorm + SiLU folded into the fp32 3x3 stencil conv to 3 channels, 128x192x64x64: one read", IN*4, None),
 "mse_search": ("scale search, 100 candidates in one pass over a 12.6 M element activation: one read (compute bound: 100 x powf per element)", 32*384*32*32*4, None),
 "qgemm2_c384": ("second-generation int8 GEMM (CTA pairs), ImageNet 32x32 conv 3x3 384->384, batch 128 (M=131072, N=384, K=3456)", None, 2*131072*384*3456),
 "qgemm2_c192": ("second-generation int8 GEMM (CTA pairs), ImageNet 64x64 conv 3x3 192->192, batch 128 (M=524288, N=192, K=1728)", None, 2*524288*192*1728),
 "qgemm2_up_c384_192": ("second-generation int8 GEMM (CTA pairs), ImageNet up-path 64x64 conv 3x3 384->192, batch 128 (M=524288, N=192, K=3456)", None, 2*524288*192*3456),
 "qgemm2_geglu_codes": ("GEGLU projection with the gate + next quantizer in the epilogue (u8 codes out), M=131072, N=3072, K=384", None, 2*131072*3072*384),
 "qgemm_i8_lin_res": ("first-generation int8 GEMM, ff.net[2] linear with residual epilogue, M=131072, N=384, K=1536 (fp32 row-major out)", None, 2*131072*384*1536),
 "qattn_imagenet": ("fused quantized attention, ImageNet self-attention T=1024, d=384, 128 (batch x head): single S accumulator, one pass per CTA", None, 4*128*1024*1024*384),
 "gemm_bf16x3_conv": ("reconstruction-loop convolution on the bf16 x 3 tcgen05 kernel (implicit GEMM, NHWC split operands, NCHW TMA store): 3x3 576->576 at 16x16, batch 32 (M=8192, N=576, K=5184); flops counted 3x (hi.hi + hi.lo + lo.hi)", None, 3*2*8192*576*5184),
 "gemm_bf16x3_wgrad": ("reconstruction-loop convolution WGRAD on the bf16 x 3 tcgen05 kernel: dW of the 3x3 576->576 conv at 16x16, batch 32 (pixels are the reduction: M=576, N=576 per tap, K=8192, 9 taps, split-K 4 with TMA reduce-add); flops counted 3x", None, 3*2*8192*576*5184),
 "gemm_bf16x3_linear": ("reconstruction-loop linear on the bf16 x 3 tcgen05 kernel: M=32768, N=3072, K=384; flops counted 3x", None, 3*2*32768*3072*384),
 "fused_adam": ("both Adam updates of a reconstruction iteration in one pass (edadm_fused_adam), ImageNet transformer-block unit: 17.5 M alphas + 20 step sizes; read g, p, m, v, write p, m, v and the cleared g", 17547284*32, None),
 "qattn_church": ("fused quantized attention, church T=1024, d=24, 800 (batch x head)", None, 4*800*1024*1024*24),
}
for name,(note,nbytes,nops) in CASES.items():
    rep=os.path.join(ROOT,"gpurun_out","r02","ncu",name+".ncu-rep")
    if not os.path.exists(rep): print("missing",name); continue
    S.summarize(rep,name,note)
    out=os.path.join(ROOT,"profiles",f"ncu_{name}_r02_summary.txt")
    txt=open(out).read()
    m=re.search(r"gpu__time_duration.sum\s+([\d.,]+)\s+(\w+)",txt)
    dur=float(m.group(1).replace(",",""))*{"us":1e-6,"usecond":1e-6,"ms":1e-3,"msecond":1e-3,"ns":1e-9,"nsecond":1e-9}.get(m.group(2),1e-6)
    extra=[]
    if nbytes: extra.append(f"algorithmic bytes {nbytes/1e6:.1f} MB / {dur*1e6:.1f} us = {nbytes/dur/1e9:.0f} GB/s = {nbytes/dur/1e9/HBM:.2f} of the measured 6546 GB/s copy bandwidth (ncu run: cold L2, serialised)")
    if nops and name.startswith("gemm_bf16x3"): extra.append(f"executed flops (3 products) {nops/1e9:.1f} GFLOP / {dur*1e6:.1f} us = {nops/dur/1e12:.0f} TFLOP/s = {nops/dur/1e12/2112.2:.2f} of the measured 2112 TFLOP/s bf16 tcgen05 burst peak (profiles/int8_peak_r02.txt)")
    elif nops: extra.append(f"algorithmic ops {nops/1e9:.1f} GOP / {dur*1e6:.1f} us = {nops/dur/1e12:.0f} TOP/s = {nops/dur/1e12/3410.6:.2f} of the measured 3411 TOP/s int8 burst peak (profiles/int8_peak_r02.txt)")
    open(out,"a").write("\n".join(extra)+"\n")
    print("   ", " | ".join(extra))
