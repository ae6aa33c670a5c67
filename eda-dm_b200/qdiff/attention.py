"""Quantized attention cores used by QuantAttnBlock / QuantBasicTransformerBlock / QuantQKMatMul.

Semantics (reference quant_block.py:128-139, 157-162, 214-233, 431-445): fake-quantize q and k, matmul, scale,
softmax in fp32, fake-quantize the probabilities (8-bit, zero-point 0) and v, matmul.
"""
import torch as th

from .quant_layer import _library_fwd


def _bmm(a, b):
    return _library_fwd(lambda x, w, _b: th.bmm(x, w), a, b, None, {})


def qk_scores_bct(q, k):
    """q, k: [b, c, t] (already quantized) -> [b, t, s]."""
    return _bmm(q.transpose(1, 2), k)


def quantized_attention_bnd(q, k, v, scale, quant_q, quant_k, quant_v, quant_w):
    """q: [b, i, d]; k, v: [b, j, d] -> [b, i, d]   (cross_attn_forward layout)."""
    sim = _bmm(quant_q(q), quant_k(k).transpose(1, 2)) * scale
    attn = sim.softmax(dim=-1)
    return _bmm(quant_w(attn), quant_v(v))


def quantized_attention_bct(q, k, v, scale, quant_q, quant_k, quant_v, quant_w):
    """q, k, v: [b, c, t] -> [b, c, t]   (DDIM AttnBlock layout; softmax over keys)."""
    qq = quant_q(q.permute(0, 2, 1))          # [b, t, c], quantized in the layout the reference uses
    kq = quant_k(k)                           # [b, c, s]
    w_ = th.softmax(_bmm(qq, kq) * scale, dim=2)
    vq = quant_v(v)
    wq = quant_w(w_.permute(0, 2, 1))         # [b, s, t]
    return _bmm(vq, wq)
