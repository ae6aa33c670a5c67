"""Torch-facing wrappers around the C ABI (include/edadm.h).

Every function takes CUDA fp32 tensors, enqueues kernels on the current torch stream and returns
torch tensors; autograd.Function classes wire the hand-written backward kernels in.  Nothing here
computes on the CPU: non-CUDA inputs raise.
"""
import math
from typing import Optional

import torch

from .native import lib, EdadmError

_partials_cache = {}
_EMPTY = torch.empty(0)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EdadmError("edadm ops need CUDA tensors (no CPU fallback for the quantized path)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _partials(device) -> torch.Tensor:
    key = (device.type, device.index)
    buf = _partials_cache.get(key)
    if buf is None:
        buf = torch.empty(lib.reduce_slots(), dtype=torch.float64, device=device)
        _partials_cache[key] = buf
    return buf


def _chan_layout(x: torch.Tensor, delta: torch.Tensor):
    """channels / inner of a per-tensor or per-dim-0 (weight, quant_layer.py:111-114) quantizer."""
    nd = delta.numel()
    if nd == 1:
        return 1, 1
    if nd != x.shape[0]:
        raise EdadmError(f"delta with {nd} elements does not match dim 0 of {tuple(x.shape)}")
    return nd, x.numel() // nd


def _qparam(t, device) -> torch.Tensor:
    """delta / zero_point as a contiguous fp32 device tensor (they are tensors / Parameters in qdiff)."""
    if not torch.is_tensor(t):
        t = torch.tensor(float(t), dtype=torch.float32, device=device)
    t = t.detach()
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    return t.contiguous().reshape(-1)


# ------------------------------------------------------------------------------------------------
# K2  UniformAffineQuantizer
# ------------------------------------------------------------------------------------------------
def uaq_forward(x, delta, zero_point, n_levels, keep_mask=None, prob=1.0, seed=0, offset=0, want_codes=False, keep_rand=None):
    _need_cuda(x)
    x = _f32c(x)
    d = _qparam(delta, x.device)
    z = _qparam(zero_point, x.device)
    channels, inner = _chan_layout(x, d)
    if z.numel() != d.numel():
        z = z.expand(d.numel()).contiguous() if z.numel() == 1 else z
    y = torch.empty_like(x)
    codes = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if want_codes else None
    if keep_mask is not None:
        keep_mask = keep_mask.to(torch.uint8).contiguous()
    if keep_rand is not None:
        keep_rand = _f32c(keep_rand)
    lib.uaq_fwd(x.data_ptr(), y.data_ptr(), _ptr(codes), d.data_ptr(), z.data_ptr(), x.numel(), channels, inner,
                int(n_levels), _ptr(keep_mask), _ptr(keep_rand), float(prob), int(seed), int(offset), _stream())
    return (y, codes) if want_codes else y


class UAQFunction(torch.autograd.Function):
    """Fake-quant with straight-through gradient (quant_layer.py:19-23, 267-274) as one kernel each way."""

    @staticmethod
    def forward(ctx, x, delta, zero_point, n_levels, keep_mask, prob, seed, offset, keep_rand=None):
        xc = _f32c(x)
        if keep_rand is not None:
            keep_rand = _f32c(keep_rand)
        y = uaq_forward(xc, delta, zero_point, n_levels, keep_mask, prob, seed, offset, keep_rand=keep_rand)
        ctx.save_for_backward(xc, delta, zero_point, keep_mask if keep_mask is not None else _EMPTY,
                              keep_rand if keep_rand is not None else _EMPTY)
        ctx.meta = (int(n_levels), float(prob), int(seed), int(offset), keep_mask is not None, keep_rand is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, delta, zero_point, keep_mask, keep_rand = ctx.saved_tensors
        n_levels, prob, seed, offset, has_mask, has_rand = ctx.meta
        gy = _f32c(gy)
        d = _qparam(delta, x.device)
        z = _qparam(zero_point, x.device)
        channels, inner = _chan_layout(x, d)
        gx = torch.empty_like(x)
        want_gd = ctx.needs_input_grad[1] and channels == 1
        gd = torch.zeros(1, dtype=torch.float32, device=x.device) if want_gd else None
        mask = keep_mask.to(torch.uint8).contiguous() if has_mask else None
        lib.uaq_bwd(gy.data_ptr(), x.data_ptr(), d.data_ptr(), z.data_ptr(), x.numel(), channels, inner, n_levels,
                    _ptr(mask), keep_rand.data_ptr() if has_rand else None, prob, seed, offset, gx.data_ptr(), _ptr(gd), 0,
                    _partials(x.device).data_ptr() if want_gd else None, _stream())
        gdelta = gd.reshape(delta.shape) if want_gd else None
        return gx, gdelta, None, None, None, None, None, None, None


def uaq_fake_quant(x, delta, zero_point, n_levels, keep_mask=None, prob=1.0, seed=0, offset=0, keep_rand=None):
    if x.numel() == 0:          # empty batch: the element-wise quantizer of nothing
        return x.clone()
    return UAQFunction.apply(x, delta, zero_point, n_levels, keep_mask, prob, seed, offset, keep_rand)


def mse_search_scores(x2d, delta, zp, n_levels, p=2.4):
    """x2d fp32 [S, inner]; delta / zp fp32 [S, K] (K <= 128 candidates per segment) -> mean |Q_k(x) - x|^p, fp32 [S, K]
    (the scores UniformAffineQuantizer.perform_1D_search minimises), all candidates in one pass over x."""
    _need_cuda(x2d)
    x2d = _f32c(x2d)
    S, inner = x2d.shape
    d = _f32c(delta.reshape(S, -1))
    z = _f32c(zp.reshape(S, -1))
    K = d.shape[1]
    scores = torch.zeros((S, K), dtype=torch.float64, device=x2d.device)
    lib.mse_search_scores(x2d.data_ptr(), S, inner, d.data_ptr(), z.data_ptr(), K, int(n_levels), float(p), scores.data_ptr(), _stream())
    return (scores / inner).float()


# ------------------------------------------------------------------------------------------------
# K3  AdaRound
# ------------------------------------------------------------------------------------------------
def adaround_init_alpha(w, delta):
    _need_cuda(w)
    w = _f32c(w)
    d = _qparam(delta, w.device)
    channels, inner = _chan_layout(w, d)
    alpha = torch.empty_like(w)
    lib.adaround_init_alpha(w.data_ptr(), d.data_ptr(), w.numel(), channels, inner, alpha.data_ptr(), _stream())
    return alpha


def adaround_forward(w, alpha, delta, zero_point, n_levels, soft, want_codes=False):
    _need_cuda(w, alpha)
    w = _f32c(w)
    a = _f32c(alpha.detach())
    d = _qparam(delta, w.device)
    z = _qparam(zero_point, w.device)
    channels, inner = _chan_layout(w, d)
    out = torch.empty_like(w)
    codes = torch.empty(w.shape, dtype=torch.uint8, device=w.device) if want_codes else None
    lib.adaround_fwd(w.data_ptr(), a.data_ptr(), d.data_ptr(), z.data_ptr(), w.numel(), channels, inner,
                     int(n_levels), 1 if soft else 0, out.data_ptr(), _ptr(codes), _stream())
    return (out, codes) if want_codes else out


class AdaRoundFunction(torch.autograd.Function):
    """W~ = (clamp(floor(W/d) + h(alpha) + zp) - zp) * d with d W~ / d alpha (adaptive_rounding.py:49-64)."""

    @staticmethod
    def forward(ctx, w, alpha, delta, zero_point, n_levels, soft):
        out = adaround_forward(w, alpha, delta, zero_point, n_levels, soft)
        ctx.save_for_backward(w, alpha, delta, zero_point)
        ctx.meta = (int(n_levels), bool(soft))
        return out

    @staticmethod
    def backward(ctx, gout):
        w, alpha, delta, zero_point = ctx.saved_tensors
        n_levels, soft = ctx.meta
        if not soft or not ctx.needs_input_grad[1]:
            return None, None, None, None, None, None
        gout = _f32c(gout)
        wc = _f32c(w)
        a = _f32c(alpha.detach())
        d = _qparam(delta, w.device)
        z = _qparam(zero_point, w.device)
        channels, inner = _chan_layout(wc, d)
        ga = torch.empty_like(a)
        lib.adaround_bwd(gout.data_ptr(), wc.data_ptr(), a.data_ptr(), d.data_ptr(), z.data_ptr(), wc.numel(), channels,
                         inner, n_levels, ga.data_ptr(), 0, _stream())
        return None, ga, None, None, None, None


def adaround_fake_quant(w, alpha, delta, zero_point, n_levels, soft):
    return AdaRoundFunction.apply(w, alpha, delta, zero_point, n_levels, soft)


class RoundRegFunction(torch.autograd.Function):
    """weight * sum(1 - |2 h(alpha) - 1|^b)  (block_recon.py:286-291)."""

    @staticmethod
    def forward(ctx, alpha, b, weight):
        a = _f32c(alpha.detach())
        loss = torch.zeros(1, dtype=torch.float32, device=a.device)
        lib.round_reg(a.data_ptr(), a.numel(), float(b), float(weight), _partials(a.device).data_ptr(), loss.data_ptr(),
                      0, None, _stream())
        ctx.save_for_backward(a)
        ctx.meta = (float(b), float(weight))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        b, weight = ctx.meta
        ga = torch.zeros_like(a)
        scratch = torch.zeros(1, dtype=torch.float32, device=a.device)
        lib.round_reg(a.data_ptr(), a.numel(), b, weight, _partials(a.device).data_ptr(), scratch.data_ptr(), 0,
                      ga.data_ptr(), _stream())
        return ga * g, None, None


def round_reg(alpha, b, weight):
    return RoundRegFunction.apply(alpha, b, weight)


# ------------------------------------------------------------------------------------------------
# K4  L_p loss
# ------------------------------------------------------------------------------------------------
class LpLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, tgt, p, per_dim1):
        _need_cuda(pred, tgt)
        pc, tc = _f32c(pred), _f32c(tgt.detach())
        n = pc.numel()
        if per_dim1:  # .sum(1).mean()
            inv_rest = pc.shape[1] / n if n else 0.0
        else:         # .mean()
            inv_rest = 1.0 / n if n else 0.0
        loss = torch.empty(1, dtype=torch.float32, device=pc.device)
        lib.lp_loss_fwd(pc.data_ptr(), tc.data_ptr(), n, float(p), float(inv_rest), _partials(pc.device).data_ptr(),
                        loss.data_ptr(), _stream())
        ctx.save_for_backward(pc, tc)
        ctx.meta = (float(p), float(inv_rest))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        pc, tc = ctx.saved_tensors
        p, inv_rest = ctx.meta
        g = _f32c(g).reshape(1)
        gp = torch.empty_like(pc)
        lib.lp_loss_bwd(pc.data_ptr(), tc.data_ptr(), pc.numel(), p, inv_rest, g.data_ptr(), gp.data_ptr(), _stream())
        return gp, None, None, None


def lp_loss(pred, tgt, p=2.0, reduction="none"):
    """Same meaning as qdiff.quant_layer.lp_loss (quant_layer.py:26-33)."""
    return LpLossFunction.apply(pred, tgt, p, reduction == "none")


def fused_adam(segments, n_segments, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, zero_grad=True):
    """One Adam step over every parameter segment (torch.optim.Adam semantics; reference qdiff/block_recon.py:113-117, 199-206).
    segments: int64 [n, 3] device table (qdiff._fused_adam.segment_table); grad / exp_avg / exp_avg_sq: flat fp32 buffers;
    lr: two device floats; step: device int64 step count (>= 1).  Parameters are updated in place through raw pointers."""
    _need_cuda(segments, grad, exp_avg, exp_avg_sq, lr, step)
    assert segments.dtype == torch.int64 and segments.is_contiguous() and step.dtype == torch.int64
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (grad, exp_avg, exp_avg_sq, lr))
    lib.fused_adam(segments.data_ptr(), int(n_segments), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                   lr.data_ptr(), step.data_ptr(), float(beta1), float(beta2), float(eps), int(bool(zero_grad)), _stream())


# ------------------------------------------------------------------------------------------------
# K1  integer path
# ------------------------------------------------------------------------------------------------
def _round_up(v, m):
    return (v + m - 1) // m * m


_BLOCK_M = 128
_MAX_BLOCK_N = 256


def gemm_block_n(n: int) -> int:
    tiles = (n + _MAX_BLOCK_N - 1) // _MAX_BLOCK_N
    bn = _round_up((n + tiles - 1) // tiles, 16)
    return max(bn, 16)


def gemm_padded_n(n: int) -> int:
    bn = gemm_block_n(n)
    return bn * ((n + bn - 1) // bn)


class PackedWeight:
    """Integer weight codes of one QuantModule K-range in the layout the GEMM consumes."""

    __slots__ = ("wq", "wsum_eff", "cw", "delta_w", "N", "Np", "R", "S", "C", "Cp", "needs_rowsum", "codes", "w4", "zoff")

    def nbytes(self):
        return self.wq.numel()


# Keep <= 4-bit weight codes nibble-packed (two per byte, half the weight bytes in HBM / L2) and let the GEMM unpack them in
# shared memory (edadm_pack_weight_w4 + edadm_qgemm_w4a8).  Bit-identical to the s8 layout; off by default because the
# unpack stage currently costs more than the halved weight traffic saves (church 32x32 conv: 79 us vs 56 us, round 1).
w4_storage = False


def pack_weight(w, delta, zero_point, n_levels, alpha=None, c_begin=0, c_end=None, want_codes=False, w4=None) -> PackedWeight:
    """w: [N, C, R, S] / [N, C, T] / [N, C] fp32 -> PackedWeight for channels [c_begin, c_end).
    w4 (default: `w4_storage` when n_levels <= 16): store 4-bit codes two per byte, [Np][R*S][Cp/2] with Cp % 32 == 0."""
    _need_cuda(w)
    w = _f32c(w.detach())
    N, Ctot = w.shape[0], w.shape[1]
    if w.dim() == 4:
        R, S = w.shape[2], w.shape[3]
    elif w.dim() == 3:
        R, S = 1, w.shape[2]
    else:
        R, S = 1, 1
    c_end = Ctot if c_end is None else c_end
    Cr = c_end - c_begin
    use_w4 = (w4_storage if w4 is None else bool(w4)) and n_levels <= 16
    Cp = _round_up(Cr, 32 if use_w4 else 16)
    Np = gemm_padded_n(N)
    d = _qparam(delta, w.device)
    z = _qparam(zero_point, w.device)
    if d.numel() == 1:
        d = d.expand(N).contiguous()
    if z.numel() == 1:
        z = z.expand(N).contiguous()
    pw = PackedWeight()
    pw.w4, pw.zoff = use_w4, None
    if use_w4:
        pw.wq = torch.empty((Np, R * S, Cp // 2), dtype=torch.uint8, device=w.device)
        wsum = torch.empty(Np, dtype=torch.int32, device=w.device)
        pw.zoff = torch.empty(Np, dtype=torch.int32, device=w.device)
        pw.cw = None
        pw.codes = torch.empty((N, Cr, R, S), dtype=torch.uint8, device=w.device) if want_codes else None
        a = None if alpha is None else _f32c(alpha.detach())
        lib.pack_weight_w4(w.data_ptr(), _ptr(a), d.data_ptr(), z.data_ptr(), N, Ctot, R, S, c_begin, c_end, Cp, Np,
                           int(n_levels), pw.wq.data_ptr(), _ptr(pw.codes), wsum.data_ptr(), pw.zoff.data_ptr(), _stream())
        pw.wsum_eff = wsum
        pw.needs_rowsum = False
        dw = torch.zeros(Np, dtype=torch.float32, device=w.device)
        dw[:N] = d
        pw.delta_w = dw
        pw.N, pw.Np, pw.R, pw.S, pw.C, pw.Cp = N, Np, R, S, Cr, Cp
        return pw
    pw.wq = torch.empty((Np, R * S, Cp), dtype=torch.int8, device=w.device)
    wsum = torch.empty(Np, dtype=torch.int32, device=w.device)
    pw.cw = torch.empty(Np, dtype=torch.int32, device=w.device)
    pw.codes = torch.empty((N, Cr, R, S), dtype=torch.uint8, device=w.device) if want_codes else None
    a = None if alpha is None else _f32c(alpha.detach())
    lib.pack_weight(w.data_ptr(), _ptr(a), d.data_ptr(), z.data_ptr(), N, Ctot, R, S, c_begin, c_end, Cp, Np,
                    int(n_levels), pw.wq.data_ptr(), _ptr(pw.codes), wsum.data_ptr(), pw.cw.data_ptr(), _stream())
    pw.wsum_eff = wsum + pw.cw * (R * S * Cr)
    pw.needs_rowsum = n_levels > 128  # resolved without a host sync: 8-bit codes may need the cw term
    dw = torch.zeros(Np, dtype=torch.float32, device=w.device)
    dw[:N] = d
    pw.delta_w = dw
    pw.N, pw.Np, pw.R, pw.S, pw.C, pw.Cp = N, Np, R, S, Cr, Cp
    return pw


class ActQuant:
    """(delta, zero_point, n_levels) of the one or two (split) activation quantizers of a QuantModule."""

    __slots__ = ("delta0", "zp0", "levels0", "split", "delta1", "zp1", "levels1", "prescale")

    def __init__(self, delta0, zp0, levels0, split=0, delta1=None, zp1=None, levels1=0, prescale=1.0):
        self.delta0, self.zp0, self.levels0 = delta0, zp0, int(levels0)
        self.split, self.delta1, self.zp1, self.levels1 = int(split), delta1, zp1, int(levels1)
        self.prescale = float(prescale)


def _batch_strided(x):
    """(tensor, batch stride) for a 4-D fp32 view whose samples are dense but spaced apart (q/k/v slices of one tensor)."""
    if x.dtype == torch.float32 and not x.is_contiguous() and x.dim() == 4:
        B, C, H, W = x.shape
        sb, sc, sh, sw = x.stride()
        if sw == 1 and (H == 1 or sh == W) and sc == H * W and sb >= C * H * W:
            return x, sb
    return _f32c(x), 0


class CatPair:
    """`th.cat([a, b], dim=1)` that was never materialised: the skip concatenation of the UNet's up path handed to a ResBlock whose
    consumers (GroupNorm statistics, the two activation producers) read the two sources in place.  Quacks like a tensor for the
    shape / device checks on the way; `materialize()` gives the real concatenation wherever a consumer needs one."""

    def __init__(self, a, b):
        assert a.shape[0] == b.shape[0] and a.shape[2:] == b.shape[2:] and a.dtype == b.dtype
        self.a, self.b = a, b
        self._full = None
        self.shape = torch.Size((a.shape[0], a.shape[1] + b.shape[1]) + tuple(a.shape[2:]))
        self.dtype, self.device, self.is_cuda, self.requires_grad = a.dtype, a.device, a.is_cuda, a.requires_grad or b.requires_grad

    def dim(self):
        return len(self.shape)

    def numel(self):
        return self.a.numel() + self.b.numel()

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def materialize(self):
        if self._full is None:
            self._full = torch.cat([self.a, self.b], dim=1)
        return self._full


def cat_slices_ok(x, split):
    """a CatPair whose two sources can be quantized into channel slices of one code tensor: boundary on a 16-channel multiple and,
    when the consumer has split quantizers, exactly at the split"""
    c0 = x.a.shape[1]
    return c0 % 16 == 0 and (not split or split == c0) and x.a.dtype == torch.float32 and x.is_cuda


def _act_quant_nhwc_cat(x: CatPair, aff, silu, aq: ActQuant, pad: int, cp: int):
    a, b = _f32c(x.a), _f32c(x.b)
    B, C0, H, W = a.shape
    C1 = b.shape[1]
    C = C0 + C1
    Cp = max(_round_up(C, 16), int(cp))
    dev = a.device
    q = torch.empty((B, H + 2 * pad, W + 2 * pad, Cp), dtype=torch.uint8, device=dev)
    d0, z0 = _qparam(aq.delta0, dev), _qparam(aq.zp0, dev)
    d1, z1, l1 = (_qparam(aq.delta1, dev), _qparam(aq.zp1, dev), aq.levels1) if aq.split else (d0, z0, aq.levels0)
    fa = fs = None
    if aff is not None:
        fa, fs = aff
    for src, coff, cs, d, z, lv in ((a, 0, C0, d0, z0, aq.levels0), (b, C0, Cp - C0, d1, z1, l1)):
        pa = None if fa is None else fa.data_ptr() + 4 * coff
        ps = None if fs is None else fs.data_ptr() + 4 * coff
        lib.act_quant_nhwc_slice(src.data_ptr(), pa, ps, C, int(silu), q.data_ptr(), Cp, coff, B, src.shape[1], H, W, cs, pad,
                                 d.data_ptr(), z.data_ptr(), lv, _stream())
    return q, None


def act_quant_nhwc(x, aq: ActQuant, pad: int, want_chsum=False, cp: int = 0):
    """x fp32 [B,C,H,W] -> u8 codes [B,H+2p,W+2p,Cp] (halo = zero-point code); Cp = cp or C rounded up to 16."""
    _need_cuda(x)
    if isinstance(x, CatPair):
        if want_chsum or aq.prescale != 1.0 or not cat_slices_ok(x, aq.split):
            x = x.materialize()
        else:
            return _act_quant_nhwc_cat(x, None, 0, aq, pad, cp)
    x, bstride = _batch_strided(x)
    B, C, H, W = x.shape
    Cp = max(_round_up(C, 16), int(cp))
    q = torch.empty((B, H + 2 * pad, W + 2 * pad, Cp), dtype=torch.uint8, device=x.device)
    chsum = torch.empty((B, H + 2 * pad, W + 2 * pad), dtype=torch.int32, device=x.device) if want_chsum else None
    dev = x.device
    d0, z0 = _qparam(aq.delta0, dev), _qparam(aq.zp0, dev)
    d1 = _qparam(aq.delta1, dev) if aq.split else None
    z1 = _qparam(aq.zp1, dev) if aq.split else None
    lib.act_quant_nhwc(x.data_ptr(), q.data_ptr(), _ptr(chsum), B, C, H, W, Cp, pad, d0.data_ptr(), z0.data_ptr(),
                       aq.levels0, aq.split, _ptr(d1), _ptr(z1), aq.levels1, aq.prescale, int(bstride), _stream())
    return q, chsum


def _cond_rows(scale, shift, B, C):
    """scale / shift [B, C, 1...] as (tensor, tensor, row stride); the two `th.chunk` halves of one embedding are used in place."""
    if scale is None:
        return None, None, 0
    sc, sh = scale.detach().reshape(B, C), shift.detach().reshape(B, C)
    if (sc.dtype == torch.float32 and sh.dtype == torch.float32 and sc.stride(1) == 1 and sh.stride(1) == 1
            and sc.stride(0) == sh.stride(0) and sc.stride(0) >= C):
        return sc, sh, sc.stride(0)
    return _f32c(sc), _f32c(sh), C


def gn_fold(x, gamma, beta, groups, eps, scale=None, shift=None):
    """GroupNorm statistics of x [B,C,...] folded into per-(sample, channel) affine (a, s), both [B, C]."""
    _need_cuda(x)
    pair = x if isinstance(x, CatPair) else None
    x = _f32c(x.a) if pair is not None else _f32c(x)
    B, C = x.shape[0], (pair.shape[1] if pair is not None else x.shape[1])
    HW = x.numel() // (B * x.shape[1])
    a = torch.empty((B, C), dtype=torch.float32, device=x.device)
    s = torch.empty((B, C), dtype=torch.float32, device=x.device)
    g = None if gamma is None else _f32c(gamma.detach())
    b = None if beta is None else _f32c(beta.detach())
    sc, sh, cond_stride = _cond_rows(scale, shift, B, C)
    if pair is not None:
        x1 = _f32c(pair.b)
        lib.gn_fold_cat(x.data_ptr(), x.shape[1], x1.data_ptr(), _ptr(g), _ptr(b), _ptr(sc), _ptr(sh), cond_stride, B, C, HW,
                        int(groups), float(eps), a.data_ptr(), s.data_ptr(), _stream())
        return a, s
    lib.gn_fold(x.data_ptr(), _ptr(g), _ptr(b), _ptr(sc), _ptr(sh), cond_stride, B, C, HW, int(groups), float(eps),
                a.data_ptr(), s.data_ptr(), _stream())
    return a, s


def norm_act_quant_nhwc(x, aff_a, aff_s, silu, aq: ActQuant, pad: int, want_chsum=False, cp: int = 0):
    """silu(a*x+s) -> u8 codes [B,H+2p,W+2p,Cp] in one pass (GroupNorm + SiLU + activation quantizer)."""
    _need_cuda(x)
    if isinstance(x, CatPair):
        if want_chsum or not cat_slices_ok(x, aq.split):
            x = x.materialize()
        else:
            return _act_quant_nhwc_cat(x, (aff_a, aff_s), silu, aq, pad, cp)
    x = _f32c(x)
    B, C, H, W = x.shape
    Cp = max(_round_up(C, 16), int(cp))
    q = torch.empty((B, H + 2 * pad, W + 2 * pad, Cp), dtype=torch.uint8, device=x.device)
    chsum = torch.empty((B, H + 2 * pad, W + 2 * pad), dtype=torch.int32, device=x.device) if want_chsum else None
    dev = x.device
    d0, z0 = _qparam(aq.delta0, dev), _qparam(aq.zp0, dev)
    d1 = _qparam(aq.delta1, dev) if aq.split else None
    z1 = _qparam(aq.zp1, dev) if aq.split else None
    lib.norm_act_quant_nhwc(x.data_ptr(), aff_a.data_ptr(), aff_s.data_ptr(), int(silu), q.data_ptr(), _ptr(chsum), B, C, H, W,
                            Cp, pad, d0.data_ptr(), z0.data_ptr(), aq.levels0, aq.split, _ptr(d1), _ptr(z1), aq.levels1, _stream())
    return q, chsum


def act_quant_rows(x2d, aq: ActQuant, want_rowsum=False, row_group=0, group_stride=0):
    """x fp32 [M,K] -> u8 codes [M,Kp].  row_group/group_stride: rows come in dense groups spaced `group_stride` elements
    apart (a [G, row_group, K] view of a larger tensor); x2d is then the storage-sharing view's base."""
    _need_cuda(x2d)
    if not row_group:
        x2d = _f32c(x2d)
    M, K = x2d.shape
    Kp = _round_up(K, 16)
    q = torch.empty((M, Kp), dtype=torch.uint8, device=x2d.device)
    rowsum = torch.empty(M, dtype=torch.int32, device=x2d.device) if want_rowsum else None
    dev = x2d.device
    d0, z0 = _qparam(aq.delta0, dev), _qparam(aq.zp0, dev)
    d1 = _qparam(aq.delta1, dev) if aq.split else None
    z1 = _qparam(aq.zp1, dev) if aq.split else None
    lib.act_quant_rows(x2d.data_ptr(), q.data_ptr(), _ptr(rowsum), M, K, Kp, d0.data_ptr(), z0.data_ptr(), aq.levels0,
                       aq.split, _ptr(d1), _ptr(z1), aq.levels1, aq.prescale, int(row_group), int(group_stride), _stream())
    return q, rowsum


def norm_act_pool2(x, aff_a, aff_s, silu):
    """avg_pool2d(silu(a*x+s), 2) in one pass; x fp32 [B,C,H,W] with even H, W."""
    _need_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    out = torch.empty((B, C, H // 2, W // 2), dtype=torch.float32, device=x.device)
    lib.norm_act_pool2(x.data_ptr(), aff_a.data_ptr(), aff_s.data_ptr(), int(silu), out.data_ptr(), B, C, H, W, _stream())
    return out


def upsample2x_codes(q_lo, C, pad, aq: ActQuant):
    """Nearest 2x upsampling of NHWC codes q_lo [B,H,W,Cp] (no halo) -> [B,2H+2p,2W+2p,Cp] with the zero-point halo ring."""
    B, H, W, Cp = q_lo.shape
    q_hi = torch.empty((B, 2 * H + 2 * pad, 2 * W + 2 * pad, Cp), dtype=torch.uint8, device=q_lo.device)
    d0, z0 = _qparam(aq.delta0, q_lo.device), _qparam(aq.zp0, q_lo.device)
    lib.upsample2x_codes(q_lo.data_ptr(), q_hi.data_ptr(), B, int(C), H, W, Cp, int(pad), d0.data_ptr(), z0.data_ptr(), aq.levels0,
                         _stream())
    return q_hi


def layernorm_quant_rows(x, norm_weight, norm_bias, eps, aq: ActQuant, want_rowsum=False):
    """LayerNorm over the last dim of x [..., K] + activation quantizer -> u8 codes [M, Kp] (one pass)."""
    _need_cuda(x)
    x2 = _f32c(x.reshape(-1, x.shape[-1]))
    M, K = x2.shape
    Kp = _round_up(K, 16)
    q = torch.empty((M, Kp), dtype=torch.uint8, device=x.device)
    rowsum = torch.empty(M, dtype=torch.int32, device=x.device) if want_rowsum else None
    g = None if norm_weight is None else _f32c(norm_weight.detach())
    b = None if norm_bias is None else _f32c(norm_bias.detach())
    d0, z0 = _qparam(aq.delta0, x.device), _qparam(aq.zp0, x.device)
    lib.layernorm_quant_rows(x2.data_ptr(), _ptr(g), _ptr(b), float(eps), q.data_ptr(), _ptr(rowsum), M, K, Kp, d0.data_ptr(),
                             z0.data_ptr(), aq.levels0, _stream())
    return q, rowsum


def layernorm_quant_rows_multi(x, norm_weight, norm_bias, eps, aqs, want_rowsum):
    """One LayerNorm pass feeding len(aqs) (<= 3) activation quantizers: list of (codes [M, Kp], rowsum or None)."""
    import ctypes
    _need_cuda(x)
    x2 = _f32c(x.reshape(-1, x.shape[-1]))
    M, K = x2.shape
    Kp = _round_up(K, 16)
    n = len(aqs)
    dev = x.device
    qs = [torch.empty((M, Kp), dtype=torch.uint8, device=dev) for _ in range(n)]
    rss = [torch.empty(M, dtype=torch.int32, device=dev) if w else None for w in want_rowsum]
    ds = [_qparam(a.delta0, dev) for a in aqs]
    zs = [_qparam(a.zp0, dev) for a in aqs]
    g = None if norm_weight is None else _f32c(norm_weight.detach())
    b = None if norm_bias is None else _f32c(norm_bias.detach())
    vp = ctypes.c_void_p * n
    lib.layernorm_quant_rows_multi(x2.data_ptr(), _ptr(g), _ptr(b), float(eps), n, vp(*[t.data_ptr() for t in qs]),
                                   vp(*[_ptr(t) for t in rss]), vp(*[t.data_ptr() for t in ds]), vp(*[t.data_ptr() for t in zs]),
                                   (ctypes.c_int * n)(*[a.levels0 for a in aqs]), M, K, Kp, _stream())
    return list(zip(qs, rss))


def geglu_quant_rows(h, aq: ActQuant, want_rowsum=False):
    """h [..., 2K] (GEGLU.proj output) -> u8 codes [M, Kp] of h[..., :K] * gelu(h[..., K:]) (one pass)."""
    _need_cuda(h)
    h2 = _f32c(h.reshape(-1, h.shape[-1]))
    M, K2 = h2.shape
    K = K2 // 2
    Kp = _round_up(K, 16)
    q = torch.empty((M, Kp), dtype=torch.uint8, device=h.device)
    rowsum = torch.empty(M, dtype=torch.int32, device=h.device) if want_rowsum else None
    d0, z0 = _qparam(aq.delta0, h.device), _qparam(aq.zp0, h.device)
    lib.geglu_quant_rows(h2.data_ptr(), q.data_ptr(), _ptr(rowsum), M, K, Kp, d0.data_ptr(), z0.data_ptr(), aq.levels0, _stream())
    return q, rowsum


def im2col_u8(q, Ho, Wo, R, S, stride):
    B, Hp, Wp, Cp = q.shape
    a = torch.empty((B * Ho * Wo, R * S * Cp), dtype=torch.uint8, device=q.device)
    lib.im2col_u8(q.data_ptr(), a.data_ptr(), B, Hp, Wp, Cp, Ho, Wo, R, S, stride, _stream())
    return a


def conv_rowsum(chsum, Ho, Wo, R, S, stride):
    B, Hp, Wp = chsum.shape
    rs = torch.empty(B * Ho * Wo, dtype=torch.int32, device=chsum.device)
    lib.conv_rowsum(chsum.data_ptr(), rs.data_ptr(), B, Hp, Wp, Ho, Wo, R, S, stride, _stream())
    return rs


gemm_profile = None   # bench.py sets this to a list to get (start_event, end_event, macs) per GEMM launch


def qgemm_i8(q, pw: PackedWeight, delta_a, zp_a, out, out_hw, bias=None, rowsum=None, a_c_offset=0,
             accumulate=False, silu=False, filter_rs=None, residual=None, bias_img=None):
    """Launch the tcgen05 GEMM.  q: [B,Hp,Wp,Cp] u8 codes (or [M,Kp] for a flat GEMM).  residual (fp32, contiguous, laid
    out like `out`) is added in the epilogue: the `x + h` of the residual / attention blocks without a separate pass."""
    if bias_img is not None:
        bias_img = _f32c(bias_img.detach().reshape(-1, pw.N))
    if residual is not None:
        if residual.dtype != torch.float32 or not residual.is_contiguous() or residual.numel() != out.numel():
            raise EdadmError("qgemm_i8: residual must be a contiguous fp32 tensor of the output's size")
    if q.dim() == 2:
        B, Hp, Wp, Cp_act = 1, 1, q.shape[0], q.shape[1]
    else:
        B, Hp, Wp, Cp_act = q.shape
    R, S = filter_rs if filter_rs is not None else (pw.R, pw.S)
    dev = q.device
    da, za = _qparam(delta_a, dev), _qparam(zp_a, dev)
    if pw.needs_rowsum and rowsum is None:
        raise EdadmError("8-bit weight codes need the activation row sums (zero-point fold); pass rowsum")
    cw = pw.cw if pw.needs_rowsum else None
    prof = gemm_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    if pw.w4:
        # 4-bit codes, two per byte; an explicit-im2col GEMM (filter_rs) sees the taps as extra channels (pitch stays % 32)
        cp_w = pw.Cp if filter_rs is None else pw.wq.shape[1] * pw.Cp // (R * S)
        lib.qgemm_w4a8(q.data_ptr(), B, Hp, Wp, Cp_act, int(a_c_offset), pw.wq.data_ptr(), pw.zoff.data_ptr(), pw.N, pw.Np,
                       R, S, cp_w, da.data_ptr(), za.data_ptr(), pw.delta_w.data_ptr(), pw.wsum_eff.data_ptr(), _ptr(bias),
                       _ptr(bias_img), _ptr(residual), out.data_ptr(), int(out_hw), 1 if accumulate else 0, int(silu), _stream())
    else:
        lib.qgemm_i8(q.data_ptr(), B, Hp, Wp, Cp_act, int(a_c_offset), pw.wq.data_ptr(), pw.N, pw.Np, R, S,
                     pw.wq.shape[2] if filter_rs is None else pw.wq.shape[1] * pw.wq.shape[2] // (R * S),
                     da.data_ptr(), za.data_ptr(), pw.delta_w.data_ptr(), pw.wsum_eff.data_ptr(), _ptr(cw), _ptr(rowsum),
                     _ptr(bias), _ptr(bias_img), _ptr(residual), out.data_ptr(), int(out_hw), 1 if accumulate else 0, int(silu), _stream())
    if prof is not None:
        ev1.record()
        m = B * (Hp - R + 1) * (Wp - S + 1)
        prof.append((ev0, ev1, m * pw.N * pw.C * pw.R * pw.S))
    return out


def qgemm_i8_split(q, pw0: PackedWeight, pw1: PackedWeight, aq0, aq1, out, out_hw, bias=None, residual=None, bias_img=None):
    """Split shortcut as one launch: the two K ranges of the code tensor q [B,Hp,Wp,Cp] (channels [0, pw0.C) / [pw0.C, ...)) against
    their own weight packs and (delta, zero_point) pairs aq0 / aq1, summed in the epilogue (falls back to two launches in C)."""
    if pw0.w4 or pw1.w4 or pw0.needs_rowsum or pw1.needs_rowsum:
        raise EdadmError("qgemm_i8_split: plain s8 weight tiles without row sums only")
    if bias_img is not None:
        bias_img = _f32c(bias_img.detach().reshape(-1, pw0.N))
    if residual is not None and (residual.dtype != torch.float32 or not residual.is_contiguous() or residual.numel() != out.numel()):
        raise EdadmError("qgemm_i8_split: residual must be a contiguous fp32 tensor of the output's size")
    if q.dim() == 2:
        B, Hp, Wp, Cp_act = 1, 1, q.shape[0], q.shape[1]
    else:
        B, Hp, Wp, Cp_act = q.shape
    dev = q.device
    d0, z0 = _qparam(aq0[0], dev), _qparam(aq0[1], dev)
    d1, z1 = _qparam(aq1[0], dev), _qparam(aq1[1], dev)
    prof = gemm_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    lib.qgemm_i8_split(q.data_ptr(), B, Hp, Wp, Cp_act, pw0.wq.data_ptr(), pw1.wq.data_ptr(), pw0.N, pw0.Np, pw0.R, pw0.S,
                       pw0.wq.shape[2], pw1.wq.shape[2], pw0.C, d0.data_ptr(), z0.data_ptr(), d1.data_ptr(), z1.data_ptr(),
                       pw0.delta_w.data_ptr(), pw1.delta_w.data_ptr(), pw0.wsum_eff.data_ptr(), pw1.wsum_eff.data_ptr(), _ptr(bias),
                       _ptr(bias_img), _ptr(residual), out.data_ptr(), int(out_hw), _stream())
    if prof is not None:
        ev1.record()
        m = B * (Hp - pw0.R + 1) * (Wp - pw0.S + 1)
        prof.append((ev0, ev1, m * pw0.N * (pw0.C + pw1.C) * pw0.R * pw0.S))
    return out


def qgemm_i8_rows_post(q, pw: PackedWeight, delta_a, zp_a, out, bias, residual, post, post_rows):
    """Linear GEMM (q [M, Kp] u8 codes -> out fp32 [M, N]) with `+ residual[m]` and then `+ post[m // post_rows]` in the epilogue."""
    if pw.w4 or pw.needs_rowsum:
        raise EdadmError("qgemm_i8_rows_post: plain s8 weight tiles without row sums only")
    if residual is None or residual.dtype != torch.float32 or not residual.is_contiguous() or residual.numel() != out.numel():
        raise EdadmError("qgemm_i8_rows_post: residual must be a contiguous fp32 tensor of the output's size")
    post = _f32c(post.detach())
    dev = q.device
    da, za = _qparam(delta_a, dev), _qparam(zp_a, dev)
    prof = gemm_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    lib.qgemm_i8_rows_post(q.data_ptr(), q.shape[0], q.shape[1], pw.wq.data_ptr(), pw.N, pw.Np, pw.wq.shape[2], da.data_ptr(),
                           za.data_ptr(), pw.delta_w.data_ptr(), pw.wsum_eff.data_ptr(), _ptr(bias), residual.data_ptr(),
                           post.data_ptr(), int(post_rows), out.data_ptr(), _stream())
    if prof is not None:
        ev1.record()
        prof.append((ev0, ev1, q.shape[0] * pw.N * pw.C * pw.R * pw.S))
    return out


def qgemm_i8_codes(q, pw: PackedWeight, delta_a, zp_a, consumer, bias=None, rowsum=None, geglu=False, want_rowsum=False):
    """Linear GEMM whose epilogue emits the u8 codes of the NEXT activation quantizer (`consumer` = (delta, zero_point,
    n_levels)) instead of fp32 -- optionally through the GEGLU gate.  q: [M, Kp] u8 codes.  Returns (codes [M, Kp_out], rowsum)."""
    if pw.w4:
        raise EdadmError("qgemm_i8_codes: nibble-packed weights are not supported by the code-emitting epilogue")
    M, Kp_act = q.shape
    dev = q.device
    n_out = pw.N // 2 if geglu else pw.N
    pitch = _round_up(n_out, 16)
    codes = torch.empty((M, pitch), dtype=torch.uint8, device=dev)
    rs_out = torch.zeros(M, dtype=torch.int32, device=dev) if want_rowsum else None
    da, za = _qparam(delta_a, dev), _qparam(zp_a, dev)
    cd, cz = _qparam(consumer[0], dev), _qparam(consumer[1], dev)
    if pw.needs_rowsum and rowsum is None:
        raise EdadmError("8-bit weight codes need the activation row sums (zero-point fold); pass rowsum")
    cw = pw.cw if pw.needs_rowsum else None
    prof = gemm_profile
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    lib.qgemm_i8_codes(q.data_ptr(), M, Kp_act, pw.wq.data_ptr(), pw.N, pw.Np, pw.wq.shape[1] * pw.wq.shape[2], da.data_ptr(),
                       za.data_ptr(), pw.delta_w.data_ptr(), pw.wsum_eff.data_ptr(), _ptr(cw), _ptr(rowsum), _ptr(bias),
                       1 if geglu else 0, cd.data_ptr(), cz.data_ptr(), int(consumer[2]), codes.data_ptr(), pitch, _ptr(rs_out),
                       _stream())
    if prof is not None:
        ev1.record()
        prof.append((ev0, ev1, M * pw.N * pw.C * pw.R * pw.S))
    return codes, rs_out


def conv3x3_small_n_ok(x, weight, kwargs):
    """True when `F.conv2d(x, weight, **kwargs)` is the narrow output-layer case edadm_conv3x3_small_n covers."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and weight.dim() == 4) or x.numel() == 0:
        return False
    if weight.shape[0] > 4 or tuple(weight.shape[2:]) != (3, 3) or x.shape[0] > 65535 or weight.shape[1] * weight.shape[0] > 3600:
        return False
    if torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad):
        return False
    kw = kwargs or {}
    one = lambda v, d: all(int(e) == d for e in (v if isinstance(v, (tuple, list)) else (v,)))
    return (one(kw.get('stride', 1), 1) and not isinstance(kw.get('padding', 0), str) and one(kw.get('padding', 0), 1)
            and one(kw.get('dilation', 1), 1) and kw.get('groups', 1) == 1)


def conv3x3_small_n(x, weight, bias=None, affine=None):
    """fp32 3x3 / stride 1 / pad 1 convolution to <= 4 output channels (the UNet output layer).  affine = (a, s, silu):
    the conv reads silu(a*x+s) with a, s [B, C] from gn_fold (GroupNorm + SiLU of the output head folded into the load)."""
    _need_cuda(x)
    x, weight = _f32c(x), _f32c(weight.detach())
    B, C, H, W = x.shape
    N = weight.shape[0]
    out = torch.empty((B, N, H, W), dtype=torch.float32, device=x.device)
    b = None if bias is None else _f32c(bias.detach())
    a_, s_, silu = (affine[0], affine[1], int(affine[2])) if affine is not None else (None, None, 0)
    lib.conv3x3_small_n(x.data_ptr(), weight.data_ptr(), _ptr(b), _ptr(a_), _ptr(s_), silu, out.data_ptr(), B, C, H, W, N, _stream())
    return out


# ------------------------------------------------------------------------------------------------
# K5  fused quantized attention
# ------------------------------------------------------------------------------------------------
class AttnQuant:
    """(delta, zero_point, n_levels) of the q, k, v and softmax quantizers of one attention site."""

    __slots__ = ("q", "k", "v", "w")

    def __init__(self, q, k, v, w):
        self.q, self.k, self.v, self.w = q, k, v, w   # each: (delta, zero_point, n_levels)


def _codes_token_major_from_bct(x, quant, prescale):
    """x fp32 [BH, d, T] -> codes [BH, T, dp] + per-token code sums [BH, T]."""
    BH, d, T = x.shape
    aq = ActQuant(quant[0], quant[1], quant[2], prescale=prescale)
    q, chsum = act_quant_nhwc(x.unsqueeze(2), aq, 0, want_chsum=True)
    return q.reshape(BH, T, q.shape[-1]), chsum.reshape(BH, T)


def _codes_rows(x2d, quant, prescale=1.0):
    aq = ActQuant(quant[0], quant[1], quant[2], prescale=prescale)
    return act_quant_rows(x2d, aq, want_rowsum=True)


def qattn(qc, kc, vc, rq, rk, rv, heads, d, Tk, aquant: AttnQuant, sm_scale, out, strides):
    """qc [BH,Tq,dp], kc [BH,Tk,dp], vc [BH,d,Tkp] u8 codes -> out (fp32, written through `strides` = (sb, sh, st, sc))."""
    BH, Tq, dp = qc.shape
    Tkp = vc.shape[-1]
    dev = qc.device
    s = [_qparam(v, dev) for v in (aquant.q[0], aquant.q[1], aquant.k[0], aquant.k[1], aquant.v[0], aquant.v[1],
                                   aquant.w[0], aquant.w[1])]
    lib.qattn_fwd(qc.data_ptr(), kc.data_ptr(), vc.data_ptr(), rq.data_ptr(), rk.data_ptr(), rv.data_ptr(), BH, int(heads),
                  Tq, int(Tk), int(d), dp, Tkp, *[t.data_ptr() for t in s], int(aquant.w[2]), float(sm_scale), out.data_ptr(),
                  int(strides[0]), int(strides[1]), int(strides[2]), int(strides[3]), _stream())
    return out


def qattn_bct(q, k, v, aquant: AttnQuant, prescale, sm_scale):
    """q, k, v fp32 [BH, d, T] (channel-major; DDIM AttnBlock and the LDM legacy attention) -> [BH, d, T]."""
    _need_cuda(q, k, v)
    BH, d, T = q.shape
    qc, rq = _codes_token_major_from_bct(q, aquant.q, prescale)
    kc, rk = _codes_token_major_from_bct(k, aquant.k, prescale)
    if v.dtype == torch.float32 and not v.is_contiguous() and v.stride(2) == 1 and v.stride(1) == T and v.stride(0) >= d * T:
        # v is a [BH, d, T] slice of the qkv tensor: quantize its rows in place (groups of d rows, stride(0) apart)
        aqv = ActQuant(aquant.v[0], aquant.v[1], aquant.v[2])
        vc, rv = act_quant_rows(v.as_strided((BH * d, T), (T, 1)), aqv, want_rowsum=True, row_group=d, group_stride=v.stride(0))
    else:
        vc, rv = _codes_rows(_f32c(v).reshape(BH * d, T), aquant.v)
    out = torch.empty((BH, d, T), dtype=torch.float32, device=q.device)
    return qattn(qc, kc, vc.reshape(BH, d, -1), rq, rk, rv.reshape(BH, d), 1, d, T, aquant, sm_scale, out, (d * T, 0, 1, T))


def qattn_bnd(q, k, v, heads, aquant: AttnQuant, sm_scale):
    """q [BH, Tq, d], k, v [BH, Tk, d] fp32 (token-major; cross_attn_forward) -> [B, Tq, heads*d] (heads merged)."""
    _need_cuda(q, k, v)
    BH, Tq, d = q.shape
    Tk = k.shape[1]
    qc, rq = _codes_rows(_f32c(q).reshape(BH * Tq, d), aquant.q)
    kc, rk = _codes_rows(_f32c(k).reshape(BH * Tk, d), aquant.k)
    aqv = ActQuant(aquant.v[0], aquant.v[1], aquant.v[2])
    vc, rv = act_quant_nhwc(_f32c(v).reshape(BH, Tk, 1, d), aqv, 0, want_chsum=True)   # "channels" = keys -> [BH,1,d,Tkp]
    B = BH // heads
    out = torch.empty((B, Tq, heads * d), dtype=torch.float32, device=q.device)
    return qattn(qc.reshape(BH, Tq, -1), kc.reshape(BH, Tk, -1), vc.reshape(BH, d, -1), rq.reshape(BH, Tq), rk.reshape(BH, Tk),
                 rv.reshape(BH, d), heads, d, Tk, aquant, sm_scale, out, (Tq * heads * d, d, heads * d, 1))


def qattn_bnd_codes(qc, rq, kc, rk, v, heads, aquant: AttnQuant, sm_scale):
    """qattn_bnd with q / k already given as u8 codes [BH, T, dp] + per-token code sums [BH, T] (emitted by the to_q / to_k
    GEMM epilogues); v fp32 [BH, Tk, d]."""
    _need_cuda(qc, kc, v)
    BH, Tq, _ = qc.shape
    Tk, d = v.shape[1], v.shape[2]
    aqv = ActQuant(aquant.v[0], aquant.v[1], aquant.v[2])
    vc, rv = act_quant_nhwc(_f32c(v).reshape(BH, Tk, 1, d), aqv, 0, want_chsum=True)   # "channels" = keys -> [BH,1,d,Tkp]
    B = BH // heads
    out = torch.empty((B, Tq, heads * d), dtype=torch.float32, device=v.device)
    return qattn(qc, kc, vc.reshape(BH, d, -1), rq, rk, rv.reshape(BH, d), heads, d, Tk, aquant, sm_scale, out,
                 (Tq * heads * d, d, heads * d, 1))


# ------------------------------------------------------------------------------------------------
# calibration path: fp32-accurate GEMM on the bf16 tensor cores (north_star (b))
# ------------------------------------------------------------------------------------------------
def split_bf16(x2d, straight=True, transposed=False):
    """x fp32 [R, C] -> (hi, lo) bf16 [R, Cp] and / or (hi_t, lo_t) bf16 [C, Rp]: x = hi + lo to ~2^-17 relative."""
    _need_cuda(x2d)
    x2d = _f32c(x2d)
    R, C = x2d.shape
    Cp, Rp = _round_up(C, 8), _round_up(R, 8)
    dev = x2d.device
    hi = torch.empty((R, Cp), dtype=torch.bfloat16, device=dev) if straight else None
    lo = torch.empty((R, Cp), dtype=torch.bfloat16, device=dev) if straight else None
    hi_t = torch.empty((C, Rp), dtype=torch.bfloat16, device=dev) if transposed else None
    lo_t = torch.empty((C, Rp), dtype=torch.bfloat16, device=dev) if transposed else None
    lib.split_bf16(x2d.data_ptr(), R, C, _ptr(hi), _ptr(lo), Cp, _ptr(hi_t), _ptr(lo_t), Rp, _stream())
    return hi, lo, hi_t, lo_t


def gemm_bf16x3(a_hi, a_lo, b_hi, b_lo, K, bias=None, splits=1):
    """out fp32 [M, N] = a . b^T (+ bias) from split operands a_* [M, Kp], b_* [N, Kp] (three tcgen05 bf16 MMAs per K step)."""
    M, Kp = a_hi.shape
    N = b_hi.shape[0]
    out = (torch.zeros if splits > 1 else torch.empty)((M, N), dtype=torch.float32, device=a_hi.device)
    lib.gemm_bf16x3(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), M, N, int(K), Kp, _ptr(bias), out.data_ptr(),
                    int(splits), _stream())
    return out


def _wgrad_splits(n_out, k_in, m_tokens):
    tiles = ((n_out + 127) // 128) * ((k_in + 127) // 128)
    return max(1, min(32, (2 * 148 + tiles - 1) // tiles, (m_tokens + 63) // 64))


class LinearBf16x3Function(torch.autograd.Function):
    """F.linear(x, w, bias) with forward, dgrad and wgrad on edadm_gemm_bf16x3 (one kernel, different operand copies)."""

    @staticmethod
    def forward(ctx, x, w, bias):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        xh, xl, xht, xlt = split_bf16(x2, True, need_dw)
        wh, wl, wht, wlt = split_bf16(w, True, need_dx)
        b = None if bias is None else _f32c(bias.detach())
        y = gemm_bf16x3(xh, xl, wh, wl, K, bias=b)
        ctx.save_for_backward(*(t for t in (xht, xlt, wht, wlt) if t is not None))
        ctx.meta = (tuple(x.shape), tuple(w.shape), need_dx, need_dw, bias is not None)
        return y.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        xshape, wshape, need_dx, need_dw, has_bias = ctx.meta
        saved = list(ctx.saved_tensors)
        xht, xlt = (saved.pop(0), saved.pop(0)) if need_dw else (None, None)
        wht, wlt = (saved.pop(0), saved.pop(0)) if need_dx else (None, None)
        N, K = wshape
        g2 = _f32c(gy.reshape(-1, N))
        M = g2.shape[0]
        gh, gl, ght, glt = split_bf16(g2, need_dx, need_dw)
        dx = dw = db = None
        if need_dx:      # dX[M, K] = dY[M, N] . (W^T[K, N])^T
            dx = gemm_bf16x3(gh, gl, wht, wlt, N).reshape(xshape)
        if need_dw:      # dW[N, K] = dY^T[N, M] . (X^T[K, M])^T  -- long reduction over the tokens: split-K
            dw = gemm_bf16x3(ght, glt, xht, xlt, M, splits=_wgrad_splits(N, K, M))
        if has_bias and ctx.needs_input_grad[2]:
            db = g2.sum(0)
        return dx, dw, db


def linear_bf16x3_ok(x, w):
    return (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and w.dim() == 2 and x.shape[-1] == w.shape[1]
            and w.shape[0] % 4 == 0 and w.shape[1] % 8 == 0 and x.numel() // x.shape[-1] >= 64)


def linear_bf16x3(x, w, bias=None):
    return LinearBf16x3Function.apply(x, w, bias)


def split_nhwc_bf16(x, pad):
    """x fp32 NCHW -> (hi, lo) bf16 NHWC [B, H + 2 pad, W + 2 pad, Cp] with a zero halo"""
    _need_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    Cp = _round_up(C, 8)
    hi = torch.empty((B, H + 2 * pad, W + 2 * pad, Cp), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    lib.split_nhwc_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), B, C, H, W, Cp, pad, _stream())
    return hi, lo


def split_filter_bf16(w, forward=True, dgrad=False):
    """w fp32 [N, C, R, S] -> forward operand (hi, lo) bf16 [N, R*S*C] (tap-major) and / or dgrad operand (hi, lo) bf16 [C, R*S*N]
    (taps reversed, channels transposed), one launch"""
    _need_cuda(w)
    w = _f32c(w)
    N, C, R, S = w.shape
    fp, dp = _round_up(R * S * C, 8), _round_up(R * S * N, 8)
    mk = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=w.device)
    fh, fl = (mk(N, fp), mk(N, fp)) if forward else (None, None)
    dh, dl = (mk(C, dp), mk(C, dp)) if dgrad else (None, None)
    lib.split_filter_bf16(w.data_ptr(), N, C, R, S, _ptr(fh), _ptr(fl), fp, _ptr(dh), _ptr(dl), dp, _stream())
    return fh, fl, dh, dl


def _conv_bf16x3_raw(x, wh, wl, n_out, R, S, c_in, pad, bias=None):
    """x fp32 NCHW, filter operand (wh, wl) bf16 [n_out, R*S*c_in] (tap-major) -> fp32 NCHW, stride 1"""
    xh, xl = split_nhwc_bf16(x, pad)
    B, Hp, Wp, Cp = xh.shape
    out = torch.empty((B, n_out, Hp - R + 1, Wp - S + 1), dtype=torch.float32, device=x.device)
    lib.conv_bf16x3(xh.data_ptr(), xl.data_ptr(), B, Hp, Wp, Cp, wh.data_ptr(), wl.data_ptr(), n_out, R, S, c_in, wh.shape[1],
                    _ptr(bias), out.data_ptr(), _stream())
    return out


conv_wgrad_on_tensor_cores = True     # mirrored from qdiff.quant_layer.backend.calib_conv_wgrad_bf16x3 by the layer


class ConvBf16x3Function(torch.autograd.Function):
    """F.conv2d(x, w, bias, stride=1, padding=(R-1)/2) with forward and dgrad on edadm_conv_bf16x3 (dgrad = the same convolution
    of dY with the flipped, transposed filter, prepared by the forward's filter split) and wgrad on edadm_conv_wgrad_bf16x3."""

    @staticmethod
    def forward(ctx, x, w, bias):
        N, C, R, S = w.shape
        need_dx = ctx.needs_input_grad[0]
        fh, fl, dh, dl = split_filter_bf16(w, True, need_dx)
        y = _conv_bf16x3_raw(x, fh, fl, N, R, S, C, (R - 1) // 2, None if bias is None else _f32c(bias.detach()))
        ctx.save_for_backward(x, w, *((dh, dl) if need_dx else ()))
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors[:2]
        N, C, R, S = w.shape
        gy = _f32c(gy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dh, dl = ctx.saved_tensors[2:]
            dx = _conv_bf16x3_raw(gy, dh, dl, C, R, S, N, (R - 1) // 2)
        if ctx.needs_input_grad[1]:
            if conv_wgrad_on_tensor_cores:
                dw = conv_wgrad_bf16x3(gy, x, R)
            else:       # library wgrad under the ambient torch.backends.cudnn.allow_tf32 (PyTorch's default: TF32)
                pad = (R - 1) // 2
                dw = torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [pad, pad], [1, 1], False, [0, 0], 1,
                                                         [False, True, False])[1]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = gy.sum((0, 2, 3))
        return dx, dw, db


def conv_wgrad_bf16x3(gy, x, R):
    """dW [N, C, R, R] of a stride-1 'same' convolution from dY [B, N, H, W] and X [B, C, H, W] on the bf16 x 3 kernel (pixels are
    the reduction dimension; few output tiles -> split-K over the pixels with TMA reduce-add)"""
    gy, x = _f32c(gy), _f32c(x)
    B, N, H, W = gy.shape
    C = x.shape[1]
    HW = H * W
    gh, gl, _, _ = split_bf16(gy.reshape(B * N, HW))
    xh = torch.empty((R, B, C, H, W), dtype=torch.bfloat16, device=x.device)      # R copies shifted along W (see edadm_conv_wgrad_bf16x3)
    xl = torch.empty_like(xh)
    lib.split_shift_bf16(x.data_ptr(), xh.data_ptr(), xl.data_ptr(), B * C * H, W, R, (R - 1) // 2, _stream())
    taps = R * R
    tiles = taps * ((N + 127) // 128) * ((C + 127) // 128)
    k_total = B * (HW // 64)
    # split-K over the pixels: enough units for 148 SMs, and at most 2048 pixels per accumulator -- the tensor core's fp32
    # accumulation truncates, its error grows linearly with the reduction length (measured 2.8e-5 at 8192 pixels)
    splits = 1
    while k_total % (splits * 2) == 0 and k_total // (splits * 2) >= 4 and (tiles * splits < 2 * 148 or k_total // splits > 32):
        splits *= 2
    out = (torch.zeros if splits > 1 else torch.empty)((taps, N, C), dtype=torch.float32, device=gy.device)
    lib.conv_wgrad_bf16x3(gh.data_ptr(), gl.data_ptr(), xh.data_ptr(), xl.data_ptr(), B, N, C, H, W, R, R, (R - 1) // 2, out.data_ptr(),
                          splits, _stream())
    return out.permute(1, 2, 0).reshape(N, C, R, R)


def conv_bf16x3_ok(x, w, kwargs):
    """stride-1 'same' convolutions whose 128-pixel tiles form a W x H x B box, both for the forward and for the dgrad"""
    if not (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and x.dim() == 4 and w.dim() == 4):
        return False
    st, pd, dl, gr = kwargs.get("stride", 1), kwargs.get("padding", 0), kwargs.get("dilation", 1), kwargs.get("groups", 1)
    as2 = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    N, C, R, S = w.shape
    if as2(st) != (1, 1) or as2(dl) != (1, 1) or gr != 1 or R != S or R % 2 == 0 or as2(pd) != ((R - 1) // 2,) * 2:
        return False
    B, _, H, W = x.shape
    if C % 8 or N % 8 or x.shape[1] != C or (B * H * W) % 128 or (H * W) % 64 or W % 8:
        return False
    hw = H * W
    if (hw % 128 if hw >= 128 else 128 % hw) or (min(hw, 128) & (min(hw, 128) - 1)):
        return False
    if W >= 128:
        return W % 128 == 0
    if 128 % W:
        return False
    rows = 128 // W
    return (H % rows == 0) if rows <= H else (rows % H == 0 and B % (rows // H) == 0)


def conv_bf16x3(x, w, bias=None):
    return ConvBf16x3Function.apply(x, w, bias)


def split_bf16_batched(x3d, straight=True, transposed=False):
    """x fp32 [G, R, C] -> (hi, lo) bf16 [G, R, Cp] and / or (hi_t, lo_t) bf16 [G, C, Rp], every matrix of the batch on its own"""
    _need_cuda(x3d)
    x3d = _f32c(x3d)
    G, R, C = x3d.shape
    Cp, Rp = _round_up(C, 8), _round_up(R, 8)
    dev = x3d.device
    mk = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)
    hi, lo = (mk(G, R, Cp), mk(G, R, Cp)) if straight else (None, None)
    hi_t, lo_t = (mk(G, C, Rp), mk(G, C, Rp)) if transposed else (None, None)
    lib.split_bf16_batched(x3d.data_ptr(), G, R, C, _ptr(hi), _ptr(lo), Cp, _ptr(hi_t), _ptr(lo_t), Rp, _stream())
    return hi, lo, hi_t, lo_t


def _bmm_nt_raw(a_hi, a_lo, b_hi, b_lo, K):
    """out fp32 [G, M, N] = a[g] . b[g]^T from split operands a_* [G, M, Kp], b_* [G, N, Kp]"""
    G, M, Kp = a_hi.shape
    N = b_hi.shape[1]
    out = torch.empty((G, M, N), dtype=torch.float32, device=a_hi.device)
    lib.gemm_bf16x3_grouped(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), G, M, N, int(K), Kp, None, out.data_ptr(),
                            1, _stream())
    return out


class BmmNTBf16x3Function(torch.autograd.Function):
    """C[g] = A[g] . B[g]^T (A [G, M, K], B [G, N, K]) with both gradients on the grouped bf16 x 3 tensor-core GEMM:
    dA = dC . B = NT(dC, B^T), dB = dC^T . A = NT(dC^T, A^T); the transposed copies come out of the same split pass."""

    @staticmethod
    def forward(ctx, a, b):
        need_da, need_db = ctx.needs_input_grad
        ah, al, aht, alt = split_bf16_batched(a, True, need_db)
        bh, bl, bht, blt = split_bf16_batched(b, True, need_da)
        ctx.save_for_backward(*(t for t in (aht, alt, bht, blt) if t is not None))
        ctx.meta = (tuple(a.shape), tuple(b.shape), need_da, need_db)
        return _bmm_nt_raw(ah, al, bh, bl, a.shape[2])

    @staticmethod
    def backward(ctx, gc):
        ashape, bshape, need_da, need_db = ctx.meta
        saved = list(ctx.saved_tensors)
        aht, alt = (saved.pop(0), saved.pop(0)) if need_db else (None, None)
        bht, blt = (saved.pop(0), saved.pop(0)) if need_da else (None, None)
        gh, gl, ght, glt = split_bf16_batched(gc, need_da, need_db)
        da = _bmm_nt_raw(gh, gl, bht, blt, bshape[1]) if need_da else None          # [G, M, K] = dC[M, N] . (B^T[K, N])^T
        db = _bmm_nt_raw(ght, glt, aht, alt, ashape[1]) if need_db else None        # [G, N, K] = dC^T[N, M] . (A^T[K, M])^T
        return da, db


def bmm_nt_bf16x3_ok(a, b):
    """shapes the grouped kernel covers: every product (forward and both gradients) needs its row count to be a multiple of 128
    and its output pitch a multiple of 4"""
    if not (a.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 3 and b.dim() == 3 and a.shape[0] == b.shape[0]):
        return False
    M, K, N = a.shape[1], a.shape[2], b.shape[1]
    return M % 128 == 0 and N % 128 == 0 and K % 4 == 0 and N % 4 == 0 and a.shape[2] == b.shape[2]


def bmm_nt_bf16x3(a, b):
    return BmmNTBf16x3Function.apply(a, b)

