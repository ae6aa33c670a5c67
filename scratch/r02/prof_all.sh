#!/bin/bash
# ncu --set full, one launch of every kernel family (round 2)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out/r02/ncu
run() { # name kernel-regex skip which
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/r02/ncu/$1 -f python scratch/r02/prof_kernels.py $4 > gpurun_out/r02/ncu/$1.log 2>&1
}
run uaq_fwd uaq_fwd_kernel 2 uaq
run uaq_bwd uaq_bwd_kernel 2 uaq
run adaround_fwd adaround_fwd_kernel 2 adaround
run adaround_bwd adaround_bwd_kernel 2 adaround
run lp_loss_fwd lp_loss_fwd 2 lp_loss
run lp_loss_bwd lp_loss_bwd 2 lp_loss
run gn_fold gn_fold_kernel 2 gn_fold
run actq_exact_silu act_quant_nhwc_tma 2 actq
run layernorm_multi layernorm_quant_rows 2 layernorm
run act_quant_rows act_quant_rows 2 rows
run conv3x3_small_n conv3x3_small_n 2 conv_small
run mse_search mse_search_kernel 2 search
run qgemm2_c384 qgemm2_kernel 2 gemm_c384
run qgemm2_c192 qgemm2_kernel 2 gemm_c192
run qgemm2_up_c384_192 qgemm2_kernel 2 gemm_up
run qgemm2_geglu_codes qgemm2_kernel 2 gemm_geglu
run qgemm_i8_lin_res qgemm_i8_kernel 2 gemm_lin
run qattn_imagenet qattn_kernel 2 attn_in
run qattn_church qattn_kernel 2 attn_church
run gemm_bf16x3_conv gemm_bf16x3_kernel 2 bf16x3_conv
run gemm_bf16x3_linear gemm_bf16x3_kernel 2 bf16x3_lin
run gemm_bf16x3_wgrad gemm_bf16x3_kernel 2 bf16x3_wgrad
ls -la gpurun_out/r02/ncu/*.ncu-rep | wc -l; du -sh gpurun_out/r02/ncu
