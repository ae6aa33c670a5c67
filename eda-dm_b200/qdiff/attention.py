"""Quantized attention cores used by QuantAttnBlock / QuantBasicTransformerBlock / the LDM legacy attention.

Semantics (reference quant_block.py:128-139, 157-162, 214-233, 431-445): fake-quantize q and k, matmul, scale,
softmax in fp32, fake-quantize the probabilities (8-bit, zero-point 0) and v, matmul.

Two routes, like QuantModule: the fused tcgen05 kernel (edadm_qattn_fwd: integer codes in, T x T matrix never
leaves the SM) whenever all four quantizers are inited and no gradient / QDrop is wanted; otherwise the fused
fake-quant kernels around library batched matmuls (reconstruction, scale search).
"""
import math
from types import MethodType

import torch as th

from edadm import ops
from .quant_layer import _library_fwd, backend


def _bmm(a, b):
    """a [G, M, K] @ b [G, K, N] on the calibration path.  Inside the reconstruction loop the product (and its two gradients) runs on
    the grouped bf16 x 3 tcgen05 GEMM whenever the shapes allow (edadm_gemm_bf16x3_grouped); otherwise library fp32."""
    if backend.calib_gemm_bf16x3 and backend.in_recon and a.is_cuda:
        bt = b.transpose(1, 2)                     # [G, N, K]: K-contiguous for k^T views, a copy otherwise (inside the split pass)
        if ops.bmm_nt_bf16x3_ok(a, bt):
            return ops.bmm_nt_bf16x3(a, bt)
    return _library_fwd(lambda x, w, _b: th.bmm(x, w), a, b, None, {})


# limits of edadm_qattn_fwd (csrc/qattn.cu): u8 codes for q/k/v and the probabilities, one CTA row of <= 4096 keys,
# blockIdx.y = batch*heads
_MAX_KEYS, _MAX_BH = 4096, 65535


def _fusable(tensors, quantizers, n_keys=0, bh=0):
    if not (backend.integer_path and backend.fused_attention) or n_keys > _MAX_KEYS or bh > _MAX_BH:
        return False
    if any(q.n_levels > 256 for q in quantizers):
        return False
    for t in tensors:
        if not t.is_cuda or t.dtype != th.float32 or (th.is_grad_enabled() and t.requires_grad):
            return False
        if t.numel() == 0:          # empty batch: nothing to launch, torch's batched matmul returns the empty result
            return False
    for q in quantizers:
        if q.inited is False or q.delta is None or q.channel_wise or (q.is_training and q.prob < 1.0):
            return False
        if th.is_grad_enabled() and q.is_training:
            return False
    return True


def _aquant(qq, qk, qv, qw):
    return ops.AttnQuant(*[(q.delta, q.zero_point, q.n_levels) for q in (qq, qk, qv, qw)])


def qk_scores_bct(q, k):
    """q, k: [b, c, t] (already quantized) -> [b, t, s]."""
    return _bmm(q.transpose(1, 2), k)


def quantized_attention_bnd(q, k, v, heads, scale, quant_q, quant_k, quant_v, quant_w):
    """q: [b*h, i, d]; k, v: [b*h, j, d] -> [b, i, h*d]   (cross_attn_forward, heads merged on the way out)."""
    if _fusable((q, k, v), (quant_q, quant_k, quant_v, quant_w), n_keys=k.shape[1], bh=q.shape[0]):
        return ops.qattn_bnd(q, k, v, heads, _aquant(quant_q, quant_k, quant_v, quant_w), scale)
    sim = _bmm(quant_q(q), quant_k(k).transpose(1, 2)) * scale
    attn = sim.softmax(dim=-1)
    out = _bmm(quant_w(attn), quant_v(v))
    bh, n, d = out.shape
    return out.reshape(bh // heads, heads, n, d).permute(0, 2, 1, 3).reshape(bh // heads, n, heads * d)


def quantized_attention_bct(q, k, v, scale, quant_q, quant_k, quant_v, quant_w):
    """q, k, v: [b, c, t] -> [b, c, t]   (DDIM AttnBlock layout; softmax over keys; scale applied to the scores)."""
    if _fusable((q, k, v), (quant_q, quant_k, quant_v, quant_w), n_keys=k.shape[2], bh=q.shape[0]):
        return ops.qattn_bct(q, k, v, _aquant(quant_q, quant_k, quant_v, quant_w), 1.0, scale)
    qq = quant_q(q.permute(0, 2, 1))          # [b, t, c], quantized in the layout the reference uses
    kq = quant_k(k)                           # [b, c, s]
    w_ = th.softmax(_bmm(qq, kq) * scale, dim=2)
    vq = quant_v(v)
    wq = quant_w(w_.permute(0, 2, 1))         # [b, s, t]
    return _bmm(vq, wq)


def legacy_attention_forward(self, qkv):
    """Replacement forward of the LDM `QKVAttentionLegacy` once its two matmul modules are the quantized ones: same
    dataflow as the reference (openaimodel.py:386-405 with QuantQKMatMul / QuantSMVMatMul), fused when possible."""
    bs, width, length = qkv.shape
    ch = width // (3 * self.n_heads)
    q, k, v = qkv.reshape(bs * self.n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    self.qkv_matmul.scale = scale
    qk, sv = self.qkv_matmul, self.smv_matmul
    # (a hook on either matmul module -- they are reconstruction units of their own, reference recon_block_Qmodel.py:48-55 --
    # must see the module's call, so the fused kernel steps aside)
    if qk.use_act_quant and sv.use_act_quant and not qk._forward_hooks and not sv._forward_hooks and \
            _fusable((qkv,), (qk.act_quantizer_q, qk.act_quantizer_k, sv.act_quantizer_v, sv.act_quantizer_w), n_keys=length,
                     bh=bs * self.n_heads):
        aq = _aquant(qk.act_quantizer_q, qk.act_quantizer_k, sv.act_quantizer_v, sv.act_quantizer_w)
        return ops.qattn_bct(q, k, v, aq, scale, 1.0).reshape(bs, self.n_heads * ch, length)
    weight = qk(q, k)
    weight = th.softmax(weight.float(), dim=-1).type(weight.dtype)
    return sv(weight, v).reshape(bs, self.n_heads * ch, length)


def patch_legacy_attention(module):
    """Called by QuantModel after swapping QKMatMul / SMVMatMul inside a QKVAttentionLegacy-like module."""
    module.forward = MethodType(legacy_attention_forward, module)
