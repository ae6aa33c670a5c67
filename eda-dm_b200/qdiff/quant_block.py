"""Quantized block wrappers with the reference's names and behaviour (qdiff/quant_block.py:20-508).

The wrappers keep the FP glue (GroupNorm, SiLU, embeddings, residuals) of the block they absorb and
route (a) every Conv/Linear through QuantModule and (b) the two attention matmuls through the
quantized-attention op.  They accept both the reference's own L0 classes (ldm.*, ddim.*, when that tree
is importable) and the from-scratch structures in unet_zoo (same attribute names).
"""
import logging
from types import MethodType

import torch as th
import torch.nn as nn

from unet_zoo import ddpm_unet as _zoo_ddpm
from unet_zoo import ldm_unet as _zoo_ldm
from .quant_layer import QuantModule, UniformAffineQuantizer, StraightThrough, backend
from . import attention as qattn

logger = logging.getLogger(__name__)


def _optional(modname, *names):
    try:
        mod = __import__(modname, fromlist=list(names))
        return [getattr(mod, n) for n in names]
    except Exception:  # the reference tree is not on sys.path (benchmarks, GPU box)
        return [None] * len(names)


_ref_ResBlock, _ref_AttentionBlock, _ref_QKMatMul, _ref_SMVMatMul, _ref_TimestepBlock = _optional(
    'ldm.modules.diffusionmodules.openaimodel', 'ResBlock', 'AttentionBlock', 'QKMatMul', 'SMVMatMul', 'TimestepBlock')
(_ref_BasicTransformerBlock,) = _optional('ldm.modules.attention', 'BasicTransformerBlock')
_ref_ResnetBlock, _ref_AttnBlock = _optional('ddim.models.diffusion', 'ResnetBlock', 'AttnBlock')

_TimestepBases = tuple(c for c in (_zoo_ldm.TimestepBlock, _ref_TimestepBlock) if c is not None)


class _Recompute(th.autograd.Function):
    """Activation recompute with the reference's semantics (ldm/modules/diffusionmodules/util.py:119-148): the forward
    runs WITHOUT recording a graph, the backward re-runs it with grad enabled and differentiates w.r.t. the inputs and
    the explicitly listed parameters.  Two consequences the reconstruction loop inherits from the reference: tensors tapped
    by forward hooks inside such a block carry no gradient (the FBR terms of checkpointed units are constants), and QDrop
    masks are re-drawn in the recompute."""

    @staticmethod
    def forward(ctx, func, n_inputs, *args):
        ctx.func = func
        ctx.inputs = list(args[:n_inputs])
        ctx.params = list(args[n_inputs:])
        with th.no_grad():
            return func(*ctx.inputs)

    @staticmethod
    def backward(ctx, *grads):
        inputs = [x.detach().requires_grad_(True) if th.is_tensor(x) else x for x in ctx.inputs]
        with th.enable_grad():
            outputs = ctx.func(*[x.view_as(x) if th.is_tensor(x) else x for x in inputs])
        wrt = [x for x in inputs if th.is_tensor(x)] + ctx.params
        got = th.autograd.grad(outputs, wrt, grads, allow_unused=True)
        it = iter(got)
        in_grads = [next(it) if th.is_tensor(x) else None for x in inputs]
        ctx.inputs = ctx.params = None
        return (None, None) + tuple(in_grads) + tuple(it)


def checkpoint(func, inputs, params, flag):
    """Recompute `func(*inputs)` in the backward pass when `flag` and a gradient is being recorded."""
    if flag and th.is_grad_enabled():
        params = [p for p in params if p.requires_grad]
        return _Recompute.apply(func, len(inputs), *(tuple(inputs) + tuple(params)))
    return func(*inputs)


def nonlinearity(x):
    return x * th.sigmoid(x)


def _prenorm_ok(conv, block):
    """GroupNorm -> SiLU -> (dropout) -> conv may be fused when the conv is a QuantModule and dropout is inactive."""
    drop = getattr(block, 'dropout', None)
    p_drop = drop.p if isinstance(drop, nn.Dropout) else (drop or 0.0)
    return isinstance(conv, QuantModule) and (not block.training or p_drop == 0)


def _emb_layers(seq, emb):
    """`emb_layers(emb)` = Linear(SiLU(emb)); every ResBlock of a forward applies the SiLU to the SAME embedding tensor, so it
    is computed once per forward and kept on the tensor object (inference only; no hooks on the SiLU)."""
    if (len(seq) == 2 and isinstance(seq[0], nn.SiLU) and not seq[0]._forward_hooks and not th.is_grad_enabled()
            and not emb.requires_grad):
        key = (emb._version, emb.data_ptr())
        cached = getattr(emb, '_edadm_silu', None)
        if cached is None or cached[0] != key:        # a reused buffer that was updated in place gets a fresh activation
            cached = (key, seq[0](emb))
            emb._edadm_silu = cached
        return seq[1](cached[1])
    return seq(emb)


def _conv_plus(conv, residual, call, post=None):
    """`residual + call()` where `call` runs the QuantModule `conv`; the add moves into the GEMM epilogue unless a hook
    observes the conv's own output (calibration caches, FBR taps).  `post` (optional, [B, 1, N]): one more row per sample,
    added after the residual -- `(call() + residual) + post` -- in the same epilogue when the layer supports it."""
    if isinstance(conv, QuantModule) and not conv._forward_hooks:
        return call(residual=residual) if post is None else call(residual=residual, post=post)
    out = call() + residual
    return out if post is None else out + post


class BaseQuantBlock(nn.Module):
    """Common state of all quantized blocks (reference quant_block.py:20-43)."""

    def __init__(self, act_quant_params: dict = {}):
        super().__init__()
        self.use_weight_quant = False
        self.use_act_quant = False
        self.can_recon = True
        self.split = 0
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)  # present but unused, as in the reference
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)


# ---- LDM / ADM residual block -------------------------------------------------------------------------
class QuantResBlock(BaseQuantBlock, *_TimestepBases):
    takes_emb = True

    def __init__(self, res, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        for name in ('channels', 'emb_channels', 'dropout', 'out_channels', 'use_conv', 'use_checkpoint',
                     'use_scale_shift_norm', 'in_layers', 'updown', 'h_upd', 'x_upd', 'emb_layers', 'out_layers',
                     'skip_connection'):
            setattr(self, name, getattr(res, name))
        self.split = 0

    def forward(self, x, emb=None, split=0):
        use_split = split != 0 and not isinstance(self.skip_connection, nn.Identity) and self.skip_connection.split == 0
        args = (x, emb, split) if use_split else (x, emb)
        return checkpoint(self._forward, args, self.parameters(), self.use_checkpoint)

    def _forward(self, x, emb, split=0):
        if emb is None:
            assert len(x) == 2
            x, emb = x
        assert x.shape[2] == x.shape[3]
        if split != 0:
            self.split = split
        return _quant_resblock_forward(self, x, emb, self.split if split != 0 else 0)

    def lazy_cat(self, h, skip, split=0):
        """The UNet's `th.cat([h, hs.pop()], dim=1)` in front of this block (openaimodel.py UNetModel.forward) as an
        `edadm.ops.CatPair` -- the concatenation is never written: GroupNorm statistics and both activation producers (in_layers'
        conv, the split skip_connection) read the two sources in place.  None whenever anything needs the real tensor (hooks of
        the calibration cache / reconstruction, gradients, fake-quant or FP state, resampling blocks, 8-bit weight row sums)."""
        from edadm import ops
        if (th.is_grad_enabled() or self._forward_hooks or self._forward_pre_hooks or self.updown or not backend.fuse_norm or not backend.lazy_cat
                or not (th.is_tensor(h) and th.is_tensor(skip) and h.is_cuda and h.dtype == th.float32 and skip.dtype == th.float32
                        and h.dim() == 4 and h.shape[0] == skip.shape[0] and h.shape[2:] == skip.shape[2:])):
            return None
        in_conv, out_conv, sc = self.in_layers[-1], self.out_layers[-1], self.skip_connection
        if not (_prenorm_ok(in_conv, self) and _prenorm_ok(out_conv, self) and isinstance(sc, QuantModule)):
            return None
        pair = ops.CatPair(h, skip)
        c0 = h.shape[1]
        if not ops.cat_slices_ok(pair, 0) or (split and split != c0) or sc.split != (split or 0) or in_conv.split:
            return None
        if sc._forward_hooks or sc._forward_pre_hooks or not sc._integer_path_ok(pair) or not in_conv.prenorm_fusable(pair, self.in_layers[0]):
            return None
        if any(p.needs_rowsum or p.w4 for m in (in_conv, sc) for p in m._packed_weights()):
            return None
        return pair


def _quant_resblock_forward(blk, x, emb, split=0):
    """Dataflow of the LDM ResBlock (reference quant_block.py:86-116) with GroupNorm + SiLU folded into the activation
    producer of the following QuantModule wherever nothing sits in between (no resampling, inactive dropout)."""
    in_conv, out_conv = blk.in_layers[-1], blk.out_layers[-1]
    fuse = _prenorm_ok(in_conv, blk) and _prenorm_ok(out_conv, blk)
    if not fuse:
        return _zoo_ldm.resblock_forward(blk, x, emb, split)
    emb_out = _emb_layers(blk.emb_layers, emb).type(x.dtype)
    while emb_out.dim() < x.dim():
        emb_out = emb_out[..., None]
    # `h + emb_out` (no scale-shift conditioning) rides on in_layers' conv as a per-(image, channel) bias unless a hook
    # observes that conv's own output
    emb_bias = emb_out if (not blk.use_scale_shift_norm and not in_conv._forward_hooks) else None
    if blk.updown:
        # in_layers[:-1] (GroupNorm, SiLU) -> h_upd -> conv: normalisation, resampling and quantization as one producer chain
        h = in_conv.forward_prenorm(x, blk.in_layers[0], resample=blk.h_upd, bias_img=emb_bias)
        x = blk.x_upd(x)
    else:
        h = in_conv.forward_prenorm(x, blk.in_layers[0], bias_img=emb_bias)
    if split and not isinstance(blk.skip_connection, nn.Identity):
        sx = blk.skip_connection(x, split=split)
    else:
        sx = blk.skip_connection(x)
    if blk.use_scale_shift_norm:
        scale, shift = th.chunk(emb_out, 2, dim=1)
        return _conv_plus(out_conv, sx, lambda **kw: out_conv.forward_prenorm(h, blk.out_layers[0], scale=scale, shift=shift, **kw))
    hin = h if emb_bias is not None else h + emb_out
    return _conv_plus(out_conv, sx, lambda **kw: out_conv.forward_prenorm(hin, blk.out_layers[0], **kw))


# ---- LDM attention: the two matmuls -----------------------------------------------------------------------
class QuantQKMatMul(BaseQuantBlock):
    def __init__(self, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.scale = None
        self.use_act_quant = False
        self.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)

    def forward(self, q, k):
        if self.use_act_quant:
            return qattn.qk_scores_bct(self.act_quantizer_q(q * self.scale), self.act_quantizer_k(k * self.scale))
        return th.einsum("bct,bcs->bts", q * self.scale, k * self.scale)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_act_quant = act_quant


class QuantSMVMatMul(BaseQuantBlock):
    def __init__(self, act_quant_params: dict = {}, sm_abit=8):
        super().__init__(act_quant_params)
        self.use_act_quant = False
        self.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
        params_w = act_quant_params.copy()
        params_w['n_bits'] = sm_abit
        params_w['symmetric'] = False
        params_w['always_zero'] = True
        self.act_quantizer_w = UniformAffineQuantizer(**params_w)

    def forward(self, weight, v):
        if self.use_act_quant:
            return th.einsum("bts,bcs->bct", self.act_quantizer_w(weight), self.act_quantizer_v(v))
        return th.einsum("bts,bcs->bct", weight, v)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_act_quant = act_quant


class QuantAttentionBlock(BaseQuantBlock):
    def __init__(self, attn, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.channels = attn.channels
        self.num_heads = attn.num_heads
        self.use_checkpoint = attn.use_checkpoint
        self.norm = attn.norm
        self.qkv = attn.qkv
        self.attention = attn.attention
        self.proj_out = attn.proj_out

    def forward(self, x):
        return checkpoint(self._forward, (x,), self.parameters(), True)

    def _forward(self, x):
        b, c, *spatial = x.shape
        x = x.flatten(2)
        if isinstance(self.qkv, QuantModule):
            qkv = self.qkv.forward_prenorm(x, self.norm, silu=False)     # GroupNorm folded into the qkv producer
        else:
            qkv = self.qkv(self.norm(x))
        a = self.attention(qkv)
        return _conv_plus(self.proj_out, x, lambda **kw: self.proj_out(a, **kw)).reshape(b, c, *spatial)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)
            elif isinstance(m, (QuantQKMatMul, QuantSMVMatMul)):
                m.set_quant_state(weight_quant, act_quant)


# ---- transformer block (ImageNet / Stable Diffusion) -------------------------------------------------------
def cross_attn_forward(self, x, context=None, mask=None, norm=None, residual=None, post=None):
    """Replacement `CrossAttention.forward` with quantized q, k, v and softmax (reference quant_block.py:204-235).
    norm / residual (optional, used by QuantBasicTransformerBlock): compute `attn(norm(x), context) + residual` with the
    LayerNorm folded into the q/k/v activation producers and the add into to_out's GEMM epilogue when nothing observes the
    intermediate tensors; otherwise module by module.  post (optional, with residual): [B, 1, C] row per sample added after the
    residual (the one-key cross-attention term of the enclosing transformer block)."""
    h = self.heads
    if mask is not None:
        raise NotImplementedError("attention masks are not used by any EDA-DM configuration")
    self_attn = context is None
    fast = _cross_attn_integer(self, x, context, norm, residual, post)
    if fast is not None:
        return fast
    if norm is not None:
        mods = (self.to_q, self.to_k, self.to_v) if self_attn else (self.to_q,)
        if all(isinstance(m, QuantModule) and m.prenorm_fusable(x, norm) for m in mods):
            q = self.to_q.forward_prenorm(x, norm, silu=False)
            if self_attn:
                k, v = self.to_k.forward_prenorm(x, norm, silu=False), self.to_v.forward_prenorm(x, norm, silu=False)
            else:
                k, v = self.to_k(context), self.to_v(context)
        else:
            xn = norm(x)
            ctx = xn if self_attn else context
            q, k, v = self.to_q(xn), self.to_k(ctx), self.to_v(ctx)
    else:
        ctx = x if self_attn else context
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
    q, k, v = (_zoo_ldm._heads_split(t, h) for t in (q, k, v))
    if self.use_act_quant:
        out = qattn.quantized_attention_bnd(q, k, v, h, self.scale, self.act_quantizer_q, self.act_quantizer_k,
                                            self.act_quantizer_v, self.act_quantizer_w)
    else:
        attn = (th.einsum('bid,bjd->bij', q, k) * self.scale).softmax(dim=-1)
        out = _zoo_ldm._heads_merge(th.einsum('bij,bjd->bid', attn, v), h)
    if residual is None:
        return self.to_out(out)
    lin, rest = self.to_out[0], self.to_out[1:]
    drop_active = any(isinstance(m, nn.Dropout) and m.p > 0 and m.training for m in rest)
    if drop_active:
        y = self.to_out(out) + residual
        return y if post is None else y + post
    return _conv_plus(lin, residual, lambda **kw: lin(out, **kw), post)


def _one_key_row(attn, x, context, norm, residual):
    """Cross attention over ONE context token (the class embedding of LDM-4 ImageNet, `ClassEmbedder` -> [B, 1, 512]): softmax
    over a single key is exactly 1 whatever q is, so every query row of a sample gets the same `to_out(Q_w(1) * Q_v(v))`;
    to_q, its LayerNorm pass and the attention matmuls drop out and to_out runs on B rows instead of B*T.  Returns that row
    [B, 1, C] (the caller adds it to the residual), or None when the shortcut does not apply."""
    if th.is_grad_enabled() or norm is None or residual is None or not attn.use_act_quant or not backend.fuse_epilogue:
        return None
    if context is None or context.shape[1] != 1:
        return None
    lin_out, rest = attn.to_out[0], attn.to_out[1:]
    if any(isinstance(m, nn.Dropout) and m.p > 0 and m.training for m in rest) or any(m._forward_hooks for m in rest):
        return None
    mods = (attn.to_q, attn.to_k, attn.to_v, lin_out)
    if not all(isinstance(m, QuantModule) for m in mods):
        return None
    qs = (attn.act_quantizer_q, attn.act_quantizer_k, attn.act_quantizer_v, attn.act_quantizer_w)
    if not qattn._fusable((x,), qs):
        return None
    if not (lin_out._integer_path_ok(x) and not lin_out._forward_hooks and not lin_out._forward_pre_hooks
            and not attn.to_q._forward_hooks and not attn.to_q._forward_pre_hooks and not norm._forward_hooks
            and lin_out._epilogue_residual(residual) is not None):
        return None
    v = attn.to_v(context)                                   # [B, 1, h*d]
    attn.to_k(context)                                       # keeps to_k's hooks / path report alive; its value cannot matter
    qw, qv = attn.act_quantizer_w, attn.act_quantizer_v
    one = th.ones(1, dtype=v.dtype, device=v.device)
    # Q_w(softmax over one key) * Q_v(v), the same row for every query -- in the arithmetic of the fused attention kernel
    # (integer product of the zero-point-free codes, scaled once by dP*dv) so both routes agree bit for bit
    p_int = qw.codes(one).to(th.int32) - qw.zero_point.to(th.int32)
    v_int = qv.codes(v).to(th.int32) - qv.zero_point.to(th.int32)
    out = (p_int * v_int).to(v.dtype) * (qw.delta.detach() * qv.delta.detach())
    row = lin_out(out)                                       # [B, 1, C]
    attn.to_q.last_path = 'elided'                           # (QuantModel.path_report)
    return row


def _cross_attn_integer(attn, x, context, norm, residual, post=None):
    """Integer-path forms of cross_attn_forward (reference quant_block.py:204-235) that never materialise fp32 q / k; None when
    they do not apply (then the module-by-module formulation below runs).

    * one context token (the class embedding of LDM-4 ImageNet, `ClassEmbedder` -> [B, 1, 512]): softmax over a single key is
      exactly 1 whatever q is, so `attn = Q_w(1)` and every query row of a sample gets the same `Q_w(1) * Q_v(v)`; to_q, its
      LayerNorm pass and the attention matmuls drop out and to_out runs on B rows instead of B*T.
    * single-head self attention: one LayerNorm pass emits the codes of to_q / to_k / to_v's activation quantizers, and the
      to_q / to_k GEMMs emit the codes of the attention's q / k quantizers from their epilogue."""
    if th.is_grad_enabled() or norm is None or residual is None or not attn.use_act_quant or not backend.fuse_epilogue:
        return None
    lin_out, rest = attn.to_out[0], attn.to_out[1:]
    if any(isinstance(m, nn.Dropout) and m.p > 0 and m.training for m in rest) or any(m._forward_hooks for m in rest):
        return None
    mods = (attn.to_q, attn.to_k, attn.to_v, lin_out)
    if not all(isinstance(m, QuantModule) for m in mods):
        return None
    qs = (attn.act_quantizer_q, attn.act_quantizer_k, attn.act_quantizer_v, attn.act_quantizer_w)
    if not qattn._fusable((x,), qs):
        return None
    if context is not None and context.shape[1] == 1:
        row = _one_key_row(attn, x, context, norm, residual)
        if row is not None:
            y = residual + row
            return y if post is None else y + post
    if context is None and attn.heads == 1 and backend.fuse_epilogue and all(m.codes_consumer_ok() for m in mods[:3]) \
            and all(m.prenorm_fusable(x, norm) for m in mods[:3]) and attn.to_q.emit_ok(attn.act_quantizer_q) \
            and attn.to_k.emit_ok(attn.act_quantizer_k):
        from edadm import ops
        lead = x.shape[:-1]
        aqs = [ops.ActQuant(m.act_quantizer.delta, m.act_quantizer.zero_point, m.act_quantizer.n_levels) for m in mods[:3]]
        (cq, rq), (ck, rk), (cv, rv) = ops.layernorm_quant_rows_multi(x, norm.weight, norm.bias, norm.eps, aqs,
                                                                      [m.needs_act_rowsum() for m in mods[:3]])
        qc, qr = attn.to_q.forward_from_codes(cq, rq, lead, emit=('plain', attn.act_quantizer_q, True))
        kc, kr = attn.to_k.forward_from_codes(ck, rk, lead, emit=('plain', attn.act_quantizer_k, True))
        v = attn.to_v.forward_from_codes(cv, rv, lead)
        B, T = x.shape[0], x.shape[1]
        out = ops.qattn_bnd_codes(qc.reshape(B, T, -1), qr.reshape(B, T), kc.reshape(B, T, -1), kr.reshape(B, T), v, 1,
                                  qattn._aquant(*qs), attn.scale)
        return _conv_plus(lin_out, residual, lambda **kw: lin_out(out, **kw), post)
    return None


def _ff_forward(ff, x, norm):
    """`ff(norm(x)) + x` of BasicTransformerBlock (ldm/modules/attention.py:  FeedForward = [GEGLU | Linear+GELU], Dropout,
    Linear) with LayerNorm, the GEGLU gate and the residual folded into the two linears' producers / epilogue."""
    first, drop, last = ff.net[0], ff.net[1], ff.net[2]
    drop_active = isinstance(drop, nn.Dropout) and drop.p > 0 and drop.training
    geglu = hasattr(first, 'proj')
    lin_in = first.proj if geglu else first[0]
    if drop_active or not isinstance(lin_in, QuantModule) or not isinstance(last, QuantModule):
        return ff(norm(x)) + x
    if (geglu and not th.is_grad_enabled() and last.codes_consumer_ok() and lin_in.prenorm_fusable(x, norm)
            and lin_in.emit_ok(last.act_quantizer, geglu=True) and last._epilogue_residual(x) is not None):
        # GEGLU gate + net[2]'s activation quantizer in the projection's GEMM epilogue: the [.., 2*inner] fp32 tensor never exists
        codes, rs = lin_in.forward_prenorm(x, norm, silu=False, emit=('geglu', last.act_quantizer, last.needs_act_rowsum()))
        return last.forward_from_codes(codes, rs, x.shape[:-1], residual=x)
    h = lin_in.forward_prenorm(x, norm, silu=False)
    if geglu:
        return _conv_plus(last, x, lambda **kw: last.forward_geglu(h, **kw))
    hh = first[1](h)
    return _conv_plus(last, x, lambda **kw: last(hh, **kw))


class QuantBasicTransformerBlock(BaseQuantBlock):
    def __init__(self, tran, act_quant_params: dict = {}, sm_abit: int = 8):
        super().__init__(act_quant_params)
        self.attn1, self.ff, self.attn2 = tran.attn1, tran.ff, tran.attn2
        self.norm1, self.norm2, self.norm3 = tran.norm1, tran.norm2, tran.norm3
        self.checkpoint = tran.checkpoint
        params_w = act_quant_params.copy()
        params_w['n_bits'] = sm_abit
        params_w['always_zero'] = True
        for attn in (self.attn1, self.attn2):
            attn.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_w = UniformAffineQuantizer(**params_w)
            attn.forward = MethodType(cross_attn_forward, attn)
            attn.use_act_quant = False

    def forward(self, x, context=None):
        return checkpoint(self._forward, (x, context), self.parameters(), self.checkpoint)

    def _forward(self, x, context=None):
        if context is None:
            assert len(x) == 2
            x, context = x
        # same dataflow as the reference (quant_block.py:254-262): x = attn1(norm1(x)) + x; x = attn2(norm2(x), ctx) + x;
        # x = ff(norm3(x)) + x -- LayerNorms, GEGLU gate and residual adds ride on the neighbouring QuantModules
        # one context token: attn2's output is one row per sample that does not depend on x -- computed first and added by
        # attn1.to_out's epilogue right after its own residual (the same two fp32 additions in the same order)
        row = _one_key_row(self.attn2, x, context, self.norm2, x) if context is not None else None
        if row is not None:
            x = self.attn1(x, norm=self.norm1, residual=x, post=row)
        else:
            x = self.attn1(x, norm=self.norm1, residual=x)
            x = self.attn2(x, context=context, norm=self.norm2, residual=x)
        return _ff_forward(self.ff, x, self.norm3)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.attn1.use_act_quant = act_quant
        self.attn2.use_act_quant = act_quant
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)


# ---- DDIM (CIFAR) blocks ----------------------------------------------------------------------------------
class QuantResnetBlock(BaseQuantBlock):
    def __init__(self, res, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.in_channels, self.out_channels = res.in_channels, res.out_channels
        self.use_conv_shortcut = res.use_conv_shortcut
        self.norm1, self.conv1, self.temb_proj = res.norm1, res.conv1, res.temb_proj
        self.norm2, self.dropout, self.conv2 = res.norm2, res.dropout, res.conv2
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                self.conv_shortcut = res.conv_shortcut
            else:
                self.nin_shortcut = res.nin_shortcut
        self.split = 0

    def forward(self, x, temb=None, split=0):
        if split != 0:
            self.split = split
        if _prenorm_ok(self.conv1, self) and _prenorm_ok(self.conv2, self):
            tb = self.temb_proj(nonlinearity(temb))[:, :, None, None]
            if self.conv1._forward_hooks:
                h = self.conv1.forward_prenorm(x, self.norm1, act_fn=nonlinearity) + tb
            else:
                h = self.conv1.forward_prenorm(x, self.norm1, act_fn=nonlinearity, bias_img=tb)
            if self.in_channels != self.out_channels:
                x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x, split=self.split)
            return _conv_plus(self.conv2, x, lambda **kw: self.conv2.forward_prenorm(h, self.norm2, act_fn=nonlinearity, **kw))
        else:
            h = self.conv1(nonlinearity(self.norm1(x)))
            h = h + self.temb_proj(nonlinearity(temb))[:, :, None, None]
            h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x, split=self.split)
        return x + h


class QuantAttnBlock(BaseQuantBlock):
    def __init__(self, attn, act_quant_params: dict = {}, sm_abit=8):
        super().__init__(act_quant_params)
        self.in_channels = attn.in_channels
        self.norm, self.q, self.k, self.v, self.proj_out = attn.norm, attn.q, attn.k, attn.v, attn.proj_out
        self.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
        params_w = act_quant_params.copy()
        params_w['n_bits'] = sm_abit
        self.act_quantizer_w = UniformAffineQuantizer(**params_w)

    def forward(self, x):
        h_ = self.norm(x)
        q, k, v = self.q(h_), self.k(h_), self.v(h_)
        b, c, h, w = q.shape
        scale = int(c) ** (-0.5)
        if self.use_act_quant:
            out = qattn.quantized_attention_bct(q.reshape(b, c, h * w), k.reshape(b, c, h * w), v.reshape(b, c, h * w), scale,
                                                self.act_quantizer_q, self.act_quantizer_k, self.act_quantizer_v,
                                                self.act_quantizer_w)
        else:
            qf = q.reshape(b, c, h * w).permute(0, 2, 1)
            w_ = th.softmax(th.bmm(qf, k.reshape(b, c, h * w)) * scale, dim=2)
            out = th.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1))
        out = out.reshape(b, c, h, w)
        return _conv_plus(self.proj_out, x, lambda **kw: self.proj_out(out, **kw))


def get_specials(quant_act=False):
    """block type -> quantized wrapper (reference quant_block.py:496-508); with `quant_act` (leaf_param) the LDM
    AttentionBlock is kept and only its two matmul modules are swapped."""
    specials = {}

    def add(zoo_cls, ref_cls, wrapper):
        specials[zoo_cls] = wrapper
        if ref_cls is not None:
            specials[ref_cls] = wrapper

    add(_zoo_ldm.ResBlock, _ref_ResBlock, QuantResBlock)
    add(_zoo_ldm.BasicTransformerBlock, _ref_BasicTransformerBlock, QuantBasicTransformerBlock)
    add(_zoo_ddpm.ResnetBlock, _ref_ResnetBlock, QuantResnetBlock)
    add(_zoo_ddpm.AttnBlock, _ref_AttnBlock, QuantAttnBlock)
    if quant_act:
        add(_zoo_ldm.QKMatMul, _ref_QKMatMul, QuantQKMatMul)
        add(_zoo_ldm.SMVMatMul, _ref_SMVMatMul, QuantSMVMatMul)
    else:
        add(_zoo_ldm.AttentionBlock, _ref_AttentionBlock, QuantAttentionBlock)
    return specials
