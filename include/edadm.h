/* libedadm.so -- C ABI of the B200-native quantized-UNet hot path of EDA-DM's `qdiff`.
 *
 * Every entry point is `extern "C"`, takes plain device pointers and sizes (no torch types), enqueues
 * on the given cudaStream_t (passed as void*) and returns 0 on success or a negative EDADM_ERR_* code;
 * edadm_last_error() returns the message of the calling thread's last failure.  There is no CPU
 * fallback: a call either runs the sm_100a kernel or fails.
 *
 * "Replaces" cites the reference (BienLuky/EDA-DM) call site a binding would swap out; the Python
 * (ctypes) binding the reference side needs is shown in INTEGRATION.md.
 *
 * Scalars that the reference keeps as tensors / nn.Parameters (activation delta, zero_point) are
 * passed as DEVICE pointers so the path never synchronises with the host.
 */
#ifndef EDADM_H
#define EDADM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDADM_OK 0
#define EDADM_ERR_ARG (-1)
#define EDADM_ERR_CUDA (-2)
#define EDADM_ERR_UNSUPPORTED (-3)

const char* edadm_last_error(void);
int edadm_abi_version(void);
/* number of double slots a `partials` workspace must hold for the two-stage reductions below */
int edadm_reduce_slots(void);

/* ---- K2: UniformAffineQuantizer fake-quant ------------------------------------------------------
 * Replaces qdiff/quant_layer.py:267-274 (UniformAffineQuantizer.forward body, incl. the QDrop
 * `torch.where(rand_like(x) < prob, x_dequant, x)`), and its autograd graph (round_ste :19-23).
 *   y = keep ? (clamp(rint(x/delta)+zp, 0, n_levels-1) - zp) * delta : x
 * delta/zero_point: 1 element (channels==1) or `channels` elements; channel of element i is
 * (i / inner) % channels.  QDrop (qdrop_prob < 1): keep_mask (u8, nullable) is an explicit mask; else keep_rand
 * (fp32, nullable) holds the uniform draws of `torch.rand_like(x)` and keep = rand < prob, i.e. the reference's own
 * random stream; else a Philox4x32-10 stream keyed by (seed, offset) draws the mask in the kernel (same draw in bwd).
 * codes (nullable) receives the integer codes as u8.                                             */
int edadm_uaq_fwd(const float* x, float* y, uint8_t* codes, const float* delta, const float* zero_point,
                  int64_t n, int64_t channels, int64_t inner, int n_levels, const uint8_t* keep_mask,
                  const float* keep_rand, float qdrop_prob, uint64_t seed, uint64_t offset, void* stream);
/* gx = straight-through gradient; gdelta (nullable, per-tensor only) = LSQ step-size gradient,
 * reduced in fp64 through `partials` (edadm_reduce_slots() doubles).                              */
int edadm_uaq_bwd(const float* gy, const float* x, const float* delta, const float* zero_point, int64_t n,
                  int64_t channels, int64_t inner, int n_levels, const uint8_t* keep_mask, const float* keep_rand,
                  float qdrop_prob, uint64_t seed, uint64_t offset, float* gx, float* gdelta, int accumulate_gdelta,
                  double* partials, void* stream);

/* ---- K3: AdaRoundQuantizer -----------------------------------------------------------------------
 * Replaces qdiff/adaptive_rounding.py:49-59 (forward, 'learned_hard_sigmoid'), :63-64
 * (get_soft_targets), :66-72 (init_alpha) and the autograd graph of alpha.                        */
int edadm_adaround_fwd(const float* w, const float* alpha, const float* delta, const float* zero_point,
                       int64_t n, int64_t channels, int64_t inner, int n_levels, int soft, float* out,
                       uint8_t* codes, void* stream);
int edadm_adaround_bwd(const float* gout, const float* w, const float* alpha, const float* delta,
                       const float* zero_point, int64_t n, int64_t channels, int64_t inner, int n_levels,
                       float* galpha, int accumulate, void* stream);
int edadm_adaround_init_alpha(const float* w, const float* delta, int64_t n, int64_t channels, int64_t inner,
                              float* alpha, void* stream);
/* Replaces the 'relaxation' branch of LossFunction.__call__, qdiff/block_recon.py:286-291:
 * loss (+)= weight * sum(1 - |2h(alpha)-1|^b); galpha (nullable) += d loss / d alpha.             */
int edadm_round_reg(const float* alpha, int64_t n, float b, float weight, double* partials, float* loss,
                    int accumulate_loss, float* galpha, void* stream);

/* ---- K4: reconstruction loss ----------------------------------------------------------------------
 * Replaces lp_loss, qdiff/quant_layer.py:26-33 as used by LossFunction (block_recon.py:271-272) and
 * the FBR per-layer terms (block_recon.py:188-191): loss = sum|pred-tgt|^p * inv_rest with
 * inv_rest = size(1)/numel  (".sum(1).mean()").                                                   */
int edadm_lp_loss_fwd(const float* pred, const float* tgt, int64_t n, float p, float inv_rest, double* partials,
                      float* loss, void* stream);
int edadm_lp_loss_bwd(const float* pred, const float* tgt, int64_t n, float p, float inv_rest, const float* gloss,
                      float* gpred, void* stream);

/* ---- K6: optimiser step of the reconstruction loop --------------------------------------------------
 * Replaces the two torch.optim.Adam(...).step() calls per iteration (qdiff/block_recon.py:113-117 construct them,
 * :199-206 step them; layer_recon.py:82-86, attn_layer_recon.py:68-72) with ONE pass over the flat gradient bucket
 * of the unit (SURVEY.md section 8b "edadm_fused_adam").  segments: device table of n_segments records
 * {float* param, int64 flat_offset, int32 count, int32 group}; lr: two device floats (group 0 = AdaRound alphas,
 * group 1 = activation step sizes); step: device int64, the 1-based step count.  Arithmetic of torch's
 * _single_tensor_adam (amsgrad off, no weight decay).  zero_grad != 0 also clears the consumed gradients.          */
int edadm_fused_adam(const void* segments, int n_segments, float* grad, float* exp_avg, float* exp_avg_sq,
                     const float* lr, const int64_t* step, double beta1, double beta2, float eps, int zero_grad,
                     void* stream);

/* ---- K1 prologue: integer-code producers ----------------------------------------------------------
 * Activation codes are exactly clamp(rint(x/delta)+zp, 0, L-1) (quant_layer.py:267-268) stored as u8.
 * act_quant_nhwc: x fp32 [B][C][H][W] -> q [B][H+2pad][W+2pad][Cp]; halo pixels hold the zero-point
 * code (zero padding of F.conv2d on dequantised values == code zp), padded channels hold 0.
 * split != 0: channels >= split use the second quantizer (quant_layer.py:415-419).
 * chsum (nullable): int32 [B][H+2pad][W+2pad] per-pixel sum of codes (for 8-bit weight zero-points).
 * prescale: x is multiplied by it (fp32) before quantization (the q*scale of QuantQKMatMul, quant_block.py:130).
 * x_batch_stride (0 = dense) / row_group + group_stride (0 = dense): read q, k, v views of one qkv tensor in place. */
int edadm_act_quant_nhwc(const float* x, uint8_t* q, int32_t* chsum, int B, int C, int H, int W, int Cp, int pad,
                         const float* delta0, const float* zp0, int n_levels0, int split, const float* delta1,
                         const float* zp1, int n_levels1, float prescale, int64_t x_batch_stride, void* stream);
int edadm_act_quant_rows(const float* x, uint8_t* q, int32_t* rowsum, int64_t M, int K, int Kp, const float* delta0,
                         const float* zp0, int n_levels0, int split, const float* delta1, const float* zp1,
                         int n_levels1, float prescale, int row_group, int64_t group_stride, void* stream);
/* f1: GroupNorm (+ scale-shift conditioning) + SiLU + quantize as one producer.  edadm_gn_fold reduces x [B][C][HW] to the
 * per-(sample, channel) affine a = rstd*gamma*(1+scale), s = (beta - mean*rstd*gamma)*(1+scale) + shift (scale/shift
 * nullable; row b of scale / shift starts at b*cond_stride, 0 = C, so the two halves of one [B][2C] embedding are
 * read in place); edadm_norm_act_quant_nhwc is edadm_act_quant_nhwc applied to silu(a*x + s).  Replaces the GroupNorm32 / SiLU
 * modules in front of a QuantModule (openaimodel.py:201-205,225-232; ddim/models/diffusion.py:120-130) on the integer path. */
int edadm_gn_fold(const float* x, const float* gamma, const float* beta, const float* scale, const float* shift,
                  int64_t cond_stride, int B, int C, int HW, int G, float eps, float* a_out, float* s_out, void* stream);
int edadm_norm_act_quant_nhwc(const float* x, const float* aff_a, const float* aff_s, int silu, uint8_t* q, int32_t* chsum,
                              int B, int C, int H, int W, int Cp, int pad, const float* delta0, const float* zp0,
                              int n_levels0, int split, const float* delta1, const float* zp1, int n_levels1, void* stream);
/* Resampling ResBlocks (openaimodel.py ResBlock up=/down=: in_layers[:-1] -> h_upd -> conv, quant_block.py:93-97).
 * edadm_norm_act_pool2: out[B][C][H/2][W/2] = avg_pool2d(silu(a*x+s), 2) in one pass (GroupNorm folded by edadm_gn_fold).
 * edadm_upsample2x_codes: nearest 2x upsampling done on the u8 codes (it commutes with the quantizer):
 * q_lo [B][H][W][Cp] -> q_hi [B][2H+2pad][2W+2pad][Cp] with the halo ring of edadm_act_quant_nhwc.                    */
int edadm_norm_act_pool2(const float* x, const float* aff_a, const float* aff_s, int silu, float* out, int B, int C, int H,
                         int W, void* stream);
int edadm_upsample2x_codes(const uint8_t* q_lo, uint8_t* q_hi, int B, int C, int H, int W, int Cp, int pad, const float* delta,
                           const float* zp, int n_levels, void* stream);

/* Transformer-block producers (ldm/modules/attention.py BasicTransformerBlock as rewritten by quant_block.py:237-262):
 * edadm_layernorm_quant_rows = nn.LayerNorm (norm1/2/3) + the activation quantizer of the linear behind it, one pass;
 * edadm_geglu_quant_rows = GEGLU's `x * F.gelu(gate)` (attention.py GEGLU.forward) + the activation quantizer of
 * FeedForward.net[2].  h is GEGLU.proj's output [M][2K]; q [M][Kp] u8; rowsum nullable.                             */
int edadm_layernorm_quant_rows(const float* x, const float* gamma, const float* beta, float eps, uint8_t* q, int32_t* rowsum,
                               int64_t M, int K, int Kp, const float* delta, const float* zp, int n_levels, void* stream);
/* The same pass feeding n (1..3) quantizers at once -- norm1 in front of to_q / to_k / to_v (qdiff/quant_block.py:254,
 * cross_attn_forward :211-213): q[t] [M][Kp], rowsum[t] (nullable array / entries), delta[t] / zp[t] device scalars.     */
int edadm_layernorm_quant_rows_multi(const float* x, const float* gamma, const float* beta, float eps, int n,
                                     uint8_t* const* q, int32_t* const* rowsum, const float* const* delta,
                                     const float* const* zp, const int* n_levels, int64_t M, int K, int Kp, void* stream);
int edadm_geglu_quant_rows(const float* h, uint8_t* q, int32_t* rowsum, int64_t M, int K, int Kp, const float* delta,
                           const float* zp, int n_levels, void* stream);
int edadm_im2col_u8(const uint8_t* q, uint8_t* a, int B, int Hp, int Wp, int Cp, int Ho, int Wo, int R, int S,
                    int stride, void* stream);
int edadm_conv_rowsum(const int32_t* chsum, int32_t* rowsum, int B, int Hp, int Wp, int Ho, int Wo, int R, int S,
                      int stride, void* stream);
/* Weight codes: nearest (UniformAffineQuantizer, quant_layer.py:267-268) or, with alpha, hard AdaRound
 * (adaptive_rounding.py:50-58).  w fp32 [N][Ctot][R][S], channel range [c_begin,c_end) ->
 * wq s8 [Np][R*S][Cp] = code - zoff (zoff = zp[n] for <=7 bit, 128 for 8 bit), wsum[n] = sum wq,
 * cw[n] = zoff - zp[n]; codes (nullable) u8 [N][c_end-c_begin][R][S].                              */
int edadm_pack_weight(const float* w, const float* alpha, const float* delta, const float* zp, int N, int Ctot, int R,
                      int S, int c_begin, int c_end, int Cp, int Np, int n_levels, int8_t* wq, uint8_t* codes,
                      int32_t* wsum, int32_t* cw, void* stream);

/* ---- K7: scale search (UniformAffineQuantizer.perform_1D_search, qdiff/quant_layer.py:150-213) ----------------------
 * One pass over x ([segments][inner]; segments = 1 per tensor, = out channels for weights) scores all K <= 128 clipping
 * candidates of every segment: scores[s*K+k] (fp64, zeroed by the caller) += sum_i |(clamp(round(x_i/d)+z, 0, L-1) - z)*d - x_i|^p
 * with (d, z) = (delta[s*K+k], zp[s*K+k]) -- the reference's lp_loss(x, Q(x), p=2.4) numerator, fp32 per element exactly as the
 * reference evaluates it (quant_layer.py:110-118, :26-33), accumulated in fp64.                                           */
int edadm_mse_search_scores(const float* x, int64_t segments, int64_t inner, const float* delta, const float* zp, int K,
                            int n_levels, float p, double* scores, void* stream);

/* ---- K1: QuantModule conv2d / conv1d / linear on integer codes (tcgen05 kind::i8) ------------------
 * Replaces `self.fwd_func(input, weight, bias, **self.fwd_kwargs)` at qdiff/quant_layer.py:434 when
 * use_weight_quant and use_act_quant are on.  Stride-1 implicit GEMM over the halo-padded NHWC codes
 * (Ho = Hp-R+1, Wo = Wp-S+1); a 2-D GEMM is B=1,Hp=1,Wp=M,R=S=1; strided convs go through
 * edadm_im2col_u8 first.  out fp32 is written as [M/out_hw][N][out_hw] (NCHW; out_hw=1 => [M][N]):
 *   out = delta_a*delta_w[n]*(acc + cw[n]*rowsum[m] - zp_a*wsum_eff[n]) + bias[n]  (+= out if accumulate)
 * then SiLU if `silu`, then + residual (nullable; fp32 laid out like out, must not alias out) -- the `x + h` / `skip_connection(x) + h` of
 * the residual and attention blocks (quant_block.py:116, :193) folded into the store.  bias_img (nullable, NCHW outputs
 * only, not together with residual): fp32 [images][N] added per (image, channel) -- the `h + emb_out` of the ResBlock
 * (quant_block.py:112-113, openaimodel.py ResBlock._forward) folded into in_layers' conv.                              */
int edadm_qgemm_i8(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, int a_c_offset, const int8_t* wq, int N, int Np,
                   int R, int S, int Cp_w, const float* delta_a, const float* zp_a, const float* delta_w,
                   const int32_t* wsum_eff, const int32_t* cw, const int32_t* rowsum, const float* bias,
                   const float* bias_img, const float* residual, float* out, int out_hw, int accumulate, int silu, void* stream);

/* Same GEMM for a linear whose ONLY consumer is the activation quantizer of the next QuantModule (the consumer's
 * `input = self.act_quantizer(input)`, qdiff/quant_layer.py:414-422): the epilogue applies that quantizer and stores its u8
 * codes, out_codes [M][out_pitch], so the fp32 tensor never exists.  geglu = 1 additionally applies GEGLU.forward
 * (ldm/modules/attention.py:37-44): codes of y[:, n] * gelu(y[:, N/2 + n]) for n < N/2.  q_rowsum (nullable, zeroed by the
 * caller) accumulates the per-row code sums the consumer GEMM / the attention kernel need for their zero-point fold.   */
int edadm_qgemm_i8_codes(const uint8_t* q, int64_t M, int Kp_act, const int8_t* wq, int N, int Np, int Cp_w,
                         const float* delta_a, const float* zp_a, const float* delta_w, const int32_t* wsum_eff,
                         const int32_t* cw, const int32_t* rowsum, const float* bias, int geglu, const float* q_delta,
                         const float* q_zp, int q_levels, uint8_t* out_codes, int out_pitch, int32_t* q_rowsum, void* stream);

/* Skip concatenation without the copy (`h = th.cat([h, hs.pop()], dim=1)` of UNetModel.forward, openaimodel.py, feeding a ResBlock):
 * edadm_gn_fold_cat = edadm_gn_fold over the virtual concatenation [x0 (C0 channels) | x1 (C - C0)] (a group may straddle both);
 * edadm_act_quant_nhwc_slice quantizes ONE source into channels [q_c_offset, q_c_offset + Cs) of the shared NHWC code tensor
 * q [B][H+2p][W+2p][q_pitch] (optional fused affine aff_* [B][aff_pitch], pre-offset to the source's first channel, + SiLU);
 * called once per source -- with the split quantizers of quant_layer.py:415-419 each source has exactly one quantizer.        */
int edadm_gn_fold_cat(const float* x0, int C0, const float* x1, const float* gamma, const float* beta, const float* scale,
                      const float* shift, int64_t cond_stride, int B, int C, int HW, int G, float eps, float* a_out, float* s_out,
                      void* stream);
int edadm_act_quant_nhwc_slice(const float* x, const float* aff_a, const float* aff_s, int aff_pitch, int silu, uint8_t* q,
                               int q_pitch, int q_c_offset, int B, int C, int H, int W, int Cs, int pad, const float* delta,
                               const float* zp, int n_levels, void* stream);

/* Split shortcut in one launch: out = conv(q[..., :split], w0) + conv(q[..., split:], w1) + bias, the two K ranges of ONE NHWC code
 * tensor q (channels [0, a_c_offset1) quantized with (delta_a0, zp_a0), the rest with (delta_a1, zp_a1)) against their own weight
 * packs, accumulated in two TMEM accumulators and combined in the epilogue with the roundings of the two-launch form
 * (edadm_qgemm_i8, then edadm_qgemm_i8 with accumulate = 1), which it falls back to where the geometry is not covered.
 * Replaces quant_layer.py:415-434 with `self.split != 0` (the skip_connection / nin_shortcut of the UNet's up path).       */
int edadm_qgemm_i8_split(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, const int8_t* wq0, const int8_t* wq1, int N, int Np, int R,
                         int S, int Cp_w0, int Cp_w1, int a_c_offset1, const float* delta_a0, const float* zp_a0, const float* delta_a1,
                         const float* zp_a1, const float* delta_w0, const float* delta_w1, const int32_t* wsum_eff0,
                         const int32_t* wsum_eff1, const float* bias, const float* bias_img, const float* residual, float* out,
                         int out_hw, void* stream);

/* Linear layer with a row-group term after the residual: out[m][n] = ((acc*scale + bias[n]) + residual[m][n]) + post[m / post_rows][n].
 * Replaces `x = attn1(norm1(x)) + x; x = attn2(norm2(x), context) + x` (quant_block.py:254-262) when the context has ONE token
 * (LDM-4 ImageNet class conditioning): softmax over one key is 1, attn2's output is one row per sample, independent of x, and
 * rides on attn1.to_out's epilogue (same fp32 additions in the same order as the two separate statements).                  */
int edadm_qgemm_i8_rows_post(const uint8_t* q, int64_t M, int Kp_act, const int8_t* wq, int N, int Np, int Cp_w, const float* delta_a,
                             const float* zp_a, const float* delta_w, const int32_t* wsum_eff, const float* bias,
                             const float* residual, const float* post, int post_rows, float* out, void* stream);

/* ---- calibration path (north_star (b)): fp32-accurate GEMM on the bf16 tensor cores ---------------------------------------
 * Replaces the fp32 library GEMM behind `self.fwd_func(input, weight, bias)` (qdiff/quant_layer.py:434) and its autograd
 * dgrad / wgrad while gradients flow (block_reconstruction, qdiff/block_recon.py:152-197).  edadm_split_bf16 writes
 * hi = bf16(x), lo = bf16(x - hi) of an fp32 matrix [rows][cols] as [rows][cols_p] and / or transposed [cols][rows_p] (pitches in
 * elements, multiples of 8, padding zeroed; pass NULL for the pair that is not needed).  edadm_gemm_bf16x3 evaluates
 *   out[M][N] = a.b^T (+ bias[n]),  a.b^T := a_hi.b_hi^T + a_hi.b_lo^T + a_lo.b_hi^T   (fp32 accumulation in TMEM)
 * from K-contiguous bf16 operands a_* [M][Kp], b_* [N][Kp]; forward, dgrad and wgrad differ only in which copies are passed
 * (see csrc/gemm_bf16x3_sm100.cu).  splits > 1: split-K, partial tiles are added into `out` (must be zeroed) by TMA reduce. */
int edadm_split_bf16(const float* x, int64_t rows, int64_t cols, void* hi, void* lo, int64_t cols_p, void* hi_t, void* lo_t,
                     int64_t rows_p, void* stream);
int edadm_gemm_bf16x3(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t M, int N, int64_t K,
                      int64_t Kp, const float* bias, float* out, int splits, void* stream);
/* Batched forms for the attention matmuls of the reconstruction loop (th.bmm / einsum of qdiff/quant_block.py:128-139, :157-162,
 * :214-233, :431-445 under autograd): x [batch][rows][cols] split per matrix; out[g] = a[g].b[g]^T for `groups` matrices stacked
 * along the rows (a_* [groups*M][Kp], b_* [groups*N][Kp], out [groups*M][N]; M % 128 == 0).                               */
int edadm_split_bf16_batched(const float* x, int64_t batch, int64_t rows, int64_t cols, void* hi, void* lo, int64_t cols_p,
                             void* hi_t, void* lo_t, int64_t rows_p, void* stream);
int edadm_gemm_bf16x3_grouped(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, int64_t groups, int64_t M,
                              int N, int64_t K, int64_t Kp, const float* bias, float* out, int splits, void* stream);

/* Convolutions of the reconstruction loop (F.conv2d behind qdiff/quant_layer.py:434 under autograd, stride 1): implicit GEMM on
 * the same bf16 x 3 kernel.  edadm_split_nhwc_bf16 writes hi / lo of x fp32 NCHW [B][C][H][W] as bf16 NHWC [B][H+2pad][W+2pad][Cp]
 * with a zero halo; edadm_conv_bf16x3 reads it through rank-4 tensor maps, one filter tap x 64 channels per K step, against
 * w_* bf16 [N][R*S*C] (tap-major, channel-minor, row pitch Kp) and stores out fp32 NCHW [B][N][Hp-R+1][Wp-S+1] (+ bias[n]).
 * dgrad is the same call on dY with the flipped, transposed filter.                                                     */
/* filter w fp32 [N][C][R][S] -> forward operand f_* bf16 [N][f_pitch >= R*S*C] (tap-major) and / or dgrad operand d_* bf16
 * [C][d_pitch >= R*S*N] (taps reversed, channels transposed); pass NULL for the pair that is not needed.              */
int edadm_split_filter_bf16(const float* w, int N, int C, int R, int S, void* f_hi, void* f_lo, int64_t f_pitch, void* d_hi, void* d_lo,
                            int64_t d_pitch, void* stream);
int edadm_split_nhwc_bf16(const float* x, void* hi, void* lo, int B, int C, int H, int W, int Cp, int pad, void* stream);
int edadm_conv_bf16x3(const void* a_hi, const void* a_lo, int B, int Hp, int Wp, int Cp, const void* w_hi, const void* w_lo, int N,
                      int R, int S, int C, int64_t Kp, const float* bias, float* out, void* stream);

/* Weight gradient of the same convolutions: dW[tap][n][c] = sum over (b, pixel) dY[b][n][pixel] * X[b][c][pixel shifted by the tap].
 * dy_* is the bf16 hi / lo split of the NCHW gradient as it lies ([B][N][H*W]; edadm_split_bf16 over the flattened rows); x_* holds S
 * copies [S][B][C][H][W] of the input pre-shifted along W by kw - pad with zero fill (edadm_split_shift_bf16: a TMA box must start
 * on a 16-byte boundary of the innermost dimension); both GEMM operands are pixel-contiguous, the row shift is a coordinate offset of
 * the rank-4 tensor map of X and the vertical zero padding is TMA's out-of-bounds fill.  out fp32 [R*S][N][C] (the caller permutes to [N][C][R][S]); splits > 1 = split-K
 * over the pixels with TMA reduce-add into the zeroed output.  Replaces autograd's convolution wgrad behind quant_layer.py:434.  */
int edadm_split_shift_bf16(const float* x, void* hi, void* lo, int64_t rows, int W, int S, int pad, void* stream);
int edadm_conv_wgrad_bf16x3(const void* dy_hi, const void* dy_lo, const void* x_hi, const void* x_lo, int B, int N, int C, int H, int W,
                            int R, int S, int pad, float* out, int splits, void* stream);

/* W4 storage: the same GEMM with the weights kept as 4-bit codes, two per byte -- wq4 u8 [Np][R*S][Cp/2] (Cp % 32 == 0; inside
 * each 32-bit word byte j = code[c0+j] | code[c0+4+j] << 4), zoff[n] = zp[n] -- and unpacked to s8 (code - zoff[n]) in
 * shared memory by dedicated warps of the GEMM kernel, tile by tile, ahead of the tensor-core MMA.  Replaces the same
 * call site as edadm_qgemm_i8 (quant_layer.py:434) for n_bits <= 4 weight quantizers.                                */
int edadm_pack_weight_w4(const float* w, const float* alpha, const float* delta, const float* zp, int N, int Ctot, int R,
                         int S, int c_begin, int c_end, int Cp, int Np, int n_levels, uint8_t* wq4, uint8_t* codes,
                         int32_t* wsum, int32_t* zoff, void* stream);
int edadm_qgemm_w4a8(const uint8_t* q, int B, int Hp, int Wp, int Cp_act, int a_c_offset, const uint8_t* wq4,
                     const int32_t* zoff, int N, int Np, int R, int S, int Cp_w, const float* delta_a, const float* zp_a,
                     const float* delta_w, const int32_t* wsum_eff, const float* bias, const float* bias_img,
                     const float* residual, float* out, int out_hw, int accumulate, int silu, void* stream);

/* fp32 3x3 convolution (stride 1, zero padding 1) with N <= 4 output channels: the UNet's output layer, whose input the
 * reference leaves un-quantized (qdiff/quant_model.py `disable_network_output_quantization`), so it runs as
 * fp32 activations x fake-quantized 8-bit weights (quant_layer.py:421-434 with disable_act_quant).  x [B][C][H][W],
 * w [N][C][3][3] (already fake-quantized), bias [N] or NULL, out [B][N][H][W].  aff_a / aff_s (nullable, [B][C], from
 * edadm_gn_fold) + silu: the conv reads silu(a*x+s) -- the `out` head's GroupNorm32 + SiLU (openaimodel.py:942-946). */
int edadm_conv3x3_small_n(const float* x, const float* w, const float* bias, const float* aff_a, const float* aff_s, int silu,
                          float* out, int B, int C, int H, int W, int N, void* stream);

/* ---- K5: fused quantized attention (tcgen05 kind::i8 for Q.K^T and P.V, softmax + P quantization on chip) ----
 * Replaces the bmm/einsum - softmax - fake-quant chain of QuantAttnBlock.forward (qdiff/quant_block.py:431-445),
 * QuantQKMatMul + softmax + QuantSMVMatMul (:128-139, :157-162, openaimodel.py:402-405) and cross_attn_forward
 * (:214-233).  qc/kc: u8 codes [BH][T][dp] (token-major), vc: u8 codes [BH][d][Tkp] (channel-major), produced by
 * edadm_act_quant_nhwc / _rows together with the code sums rq[BH][Tq], rk[BH][Tk], rv[BH][d].
 *   out[b,h,t,c] = dP*dv * sum_s (Pq[t,s]-zP)(v[c,s]-zv),  Pq = clamp(rint(softmax_s(dq*dk*scale*sum_c(q-zq)(k-zk))/dP)+zP)
 * written at b*o_sb + h*o_sh + t*o_st + c*o_sc (bh = b*heads + h).                                                   */
int edadm_qattn_fwd(const uint8_t* qc, const uint8_t* kc, const uint8_t* vc, const int32_t* rq, const int32_t* rk,
                    const int32_t* rv, int BH, int heads, int Tq, int Tk, int d, int dp, int Tkp, const float* dq,
                    const float* zq, const float* dk, const float* zk, const float* dv, const float* zv, const float* dpq,
                    const float* zpq, int p_levels, float sm_scale, float* out, int64_t o_sb, int64_t o_sh, int64_t o_st,
                    int64_t o_sc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDADM_H */
