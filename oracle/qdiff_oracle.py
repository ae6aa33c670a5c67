"""CPU restatement of the reference fake-quant algorithm (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Plain fp32 torch ops on whatever device the inputs live on (CPU in the unit tests), each function
citing the reference lines it restates.  Parity is PINNED: `oracle/make_golden.py` runs the
unmodified reference (imported from /root/reference in the build container) on seeded inputs and
commits its outputs under tests/golden/; tests/test_oracle_golden.py checks every function here
against those vectors.  The reference itself ships no tests or golden vectors.
"""
import math
import random

import torch
import torch.nn.functional as F


# ---- qdiff/quant_layer.py:19-23 ------------------------------------------------------------------
def round_ste(x):
    return (x.round() - x).detach() + x


# ---- qdiff/quant_layer.py:26-33 ------------------------------------------------------------------
def lp_loss(pred, tgt, p=2.0, reduction="none"):
    if reduction == "none":
        return (pred - tgt).abs().pow(p).sum(1).mean()
    return (pred - tgt).abs().pow(p).mean()


# ---- qdiff/quant_layer.py:267-268: the integer codes the reference never materialises --------------
def uaq_codes(x, delta, zero_point, n_levels):
    x_int = torch.round(x / delta) + zero_point
    return torch.clamp(x_int, 0, n_levels - 1)


# ---- qdiff/quant_layer.py:267-276 (forward body; `keep` replaces rand_like(x) < prob) ---------------
def uaq_forward(x, delta, zero_point, n_levels, keep=None, **_unused):
    x_int = round_ste(x / delta) + zero_point
    x_quant = torch.clamp(x_int, 0, n_levels - 1)
    x_dequant = (x_quant - zero_point) * delta
    if keep is not None:
        return torch.where(keep, x_dequant, x)
    return x_dequant


# ---- qdiff/quant_layer.py:95-105 ------------------------------------------------------------------
def calculate_qparams(min_val, max_val, n_levels, eps=1e-8):
    quant_min, quant_max = 0, n_levels - 1
    min_val_neg = torch.min(min_val, torch.zeros_like(min_val))
    max_val_pos = torch.max(max_val, torch.zeros_like(max_val))
    scale = (max_val_pos - min_val_neg) / float(quant_max - quant_min)
    scale = torch.max(scale, torch.tensor(eps, dtype=torch.float32))
    zero_point = quant_min - torch.round(min_val_neg / scale)
    zero_point = torch.clamp(zero_point, quant_min, quant_max)
    return scale, zero_point


# ---- qdiff/quant_layer.py:108-118 -----------------------------------------------------------------
def _quantize_minmax(x, x_max, x_min, n_levels, channel_wise):
    delta, zero_point = calculate_qparams(x_min, x_max, n_levels)
    if channel_wise:
        shape = [1] * x.dim()
        shape[0] = x.shape[0]
        delta, zero_point = delta.reshape(shape), zero_point.reshape(shape)
    x_int = torch.round(x / delta)
    x_quant = torch.clamp(x_int + zero_point, 0, n_levels - 1)
    return (x_quant - zero_point) * delta


def _search_score(pred, tgt, channel_wise):
    x = (pred - tgt).abs().pow(2.4)  # quant_layer.py:87-93
    return x.mean() if not channel_wise else torch.flatten(x, 1).mean(1)


# ---- qdiff/quant_layer.py:150-213 -----------------------------------------------------------------
def search_1d(x, n_levels, channel_wise, one_side_dist, num=100):
    if channel_wise:
        y = torch.flatten(x, 1)
        x_min, x_max = y.amin(1), y.amax(1)
    else:
        x_min, x_max = x.amin(), x.amax()
    xrange = torch.max(x_min.abs(), x_max)
    if not channel_wise:  # :165-199 batched per-tensor search
        thres = xrange / num * torch.arange(1, num + 1, device=x.device)
        new_min = torch.zeros_like(thres) if one_side_dist == "pos" else -thres
        new_max = torch.zeros_like(thres) if one_side_dist == "neg" else thres
        scale = (new_max - new_min) / float(n_levels - 1)
        scale = torch.max(scale, torch.tensor(1e-8, dtype=torch.float32))
        zero_point = -torch.round(new_min / scale)
        zero_point = torch.clamp(zero_point, 0, n_levels - 1).view(-1, 1)
        scale = scale.view(-1, 1)
        scores = []
        for i in range(0, num, 8):
            x_int = (x.reshape(1, -1) / scale[i:i + 8]).round()
            x_int = torch.max(torch.min(x_int, n_levels - 1 - zero_point[i:i + 8]), -zero_point[i:i + 8])
            x_sim = x_int * scale[i:i + 8]
            scores.append((x_sim - x.reshape(1, -1)).abs().pow(2.4).mean(1))
        ind = torch.argmin(torch.hstack(scores))
        return new_min[ind], new_max[ind]
    best_score = torch.zeros_like(x_min) + 1e10
    best_min, best_max = x_min.clone(), x_max.clone()
    for i in range(1, num + 1):  # :201-213
        thres = xrange / num * i
        new_min = torch.zeros_like(x_min) if one_side_dist == "pos" else -thres
        new_max = torch.zeros_like(x_max) if one_side_dist == "neg" else thres
        x_q = _quantize_minmax(x, new_max, new_min, n_levels, channel_wise)
        score = _search_score(x, x_q, channel_wise)
        best_min = torch.where(score < best_score, new_min, best_min)
        best_max = torch.where(score < best_score, new_max, best_max)
        best_score = torch.min(score, best_score)
    return best_min, best_max


# ---- qdiff/quant_layer.py:120-147 -----------------------------------------------------------------
def search_2d(x, n_bits, channel_wise, num=100):
    """asymmetric two-sided range: every clipping width xrange * i / num (i = 1..num) with every zero-point 0..n_levels-1"""
    n_levels = 2 ** n_bits
    if channel_wise:
        y = torch.flatten(x, 1)
        x_min, x_max = y.amin(1), y.amax(1)
        x_max = torch.max(x_max, torch.zeros_like(x_max))   # :124-126 one-sided channels
        x_min = torch.min(x_min, torch.zeros_like(x_min))
    else:
        x_min, x_max = x.amin(), x.amax()
    xrange = x_max - x_min
    best_score = torch.zeros_like(x_min) + 1e10
    best_min, best_max = x_min.clone(), x_max.clone()
    for i in range(1, num + 1):
        tmp_min = torch.zeros_like(x_min)
        tmp_max = xrange / num * i
        tmp_delta = (tmp_max - tmp_min) / (2 ** n_bits - 1)
        for zp in range(0, n_levels):
            new_min, new_max = tmp_min - zp * tmp_delta, tmp_max - zp * tmp_delta
            x_q = _quantize_minmax(x, new_max, new_min, n_levels, channel_wise)
            score = _search_score(x, x_q, channel_wise)
            best_min = torch.where(score < best_score, new_min, best_min)
            best_max = torch.where(score < best_score, new_max, best_max)
            best_score = torch.min(best_score, score)
    return best_min, best_max


# ---- qdiff/quant_layer.py:215-244 (init path for scale_method='mse', sym or one-sided) --------------
def init_scale(x, n_bits, channel_wise, sym=True, running=None):
    """Returns (delta, zero_point, one_side_dist, running) following get_x_min_x_max +
    update_quantize_range (:79-85, only when `running` is a dict == leaf_param) + calculate_qparams."""
    n_levels = 2 ** n_bits
    one_side = "pos" if x.min() >= 0.0 else "neg" if x.max() <= 0.0 else "no"
    if one_side != "no" or sym:          # :227-231
        best_min, best_max = search_1d(x, n_levels, channel_wise, one_side)
    else:
        best_min, best_max = search_2d(x, n_bits, channel_wise)
    if running is not None:
        if running.get("min") is None:
            running["min"], running["max"] = best_min, best_max
        running["min"] = 0.1 * best_min + 0.9 * running["min"]
        running["max"] = 0.1 * best_max + 0.9 * running["max"]
        best_min, best_max = running["min"], running["max"]
    delta, zp = calculate_qparams(best_min, best_max, n_levels)
    if channel_wise:
        shape = [1] * x.dim()
        shape[0] = x.shape[0]
        delta, zp = delta.reshape(shape), zp.reshape(shape)
    return delta, zp, one_side


# ---- qdiff/quant_layer.py:278-345 (scale_method 'max' / 'max_scale': the tensor's extrema, no search) ----------
def init_scale_max(x, n_bits, channel_wise, sym=False, always_zero=False, scale_method="max"):
    """(delta, zero_point) of init_quantization_scale_2.  Per tensor (:303-327): Python-float (double) arithmetic on
    `.item()` extrema -- x_min = min(min, 0), x_max = max(max, 0), both times (n_bits + 2) / 8 for 'max_scale';
    delta = max(|x_min|, x_max) / n_levels when symmetric, else (max - min) / (n_levels - 1) from the RAW extrema
    (:318, unscaled and not clamped to 0); floor 1e-8; zero_point = round(-x_min / delta) (Python round == half to even)
    unless symmetric / always_zero (0); delta is cast to x's dtype at the end.  Channel-wise (:280-302): the same per
    slice x[c], written into fp32 vectors shaped [C, 1, ...]."""
    n_levels = 2 ** n_bits
    if channel_wise:
        pairs = [init_scale_max(x[c], n_bits, False, sym, always_zero, scale_method) for c in range(x.shape[0])]
        shape = [x.shape[0]] + [1] * (x.dim() - 1)
        return (torch.stack([d for d, _ in pairs]).reshape(shape), torch.stack([z for _, z in pairs]).reshape(shape))
    lo, hi = x.min().item(), x.max().item()
    x_min, x_max = min(lo, 0), max(hi, 0)
    if "scale" in scale_method:
        x_min, x_max = x_min * (n_bits + 2) / 8, x_max * (n_bits + 2) / 8
    delta = max(abs(x_min), x_max) / n_levels if sym else float(hi - lo) / (n_levels - 1)
    if delta < 1e-8:
        delta = 1e-8
    zp = round(-x_min / delta) if not (sym or always_zero) else 0
    return torch.tensor(delta).type_as(x), torch.tensor(float(zp)).type_as(x)


# ---- qdiff/adaptive_rounding.py -------------------------------------------------------------------
GAMMA, ZETA = -0.1, 1.1


def adaround_init_alpha(w, delta):  # :66-72
    x_floor = torch.floor(w / delta)
    rest = (w / delta) - x_floor
    return -torch.log((ZETA - GAMMA) / (rest - GAMMA) - 1)


def soft_targets(alpha):  # :63-64
    return torch.clamp(torch.sigmoid(alpha) * (ZETA - GAMMA) + GAMMA, 0, 1)


def adaround_forward(w, alpha, delta, zero_point, n_levels, soft):  # :49-59
    x_floor = torch.floor(w / delta)
    x_int = x_floor + (soft_targets(alpha) if soft else (alpha >= 0).float())
    x_quant = torch.clamp(x_int + zero_point, 0, n_levels - 1)
    return (x_quant - zero_point) * delta


def adaround_codes(w, alpha, delta, zero_point, n_levels):
    return torch.clamp(torch.floor(w / delta) + (alpha >= 0).float() + zero_point, 0, n_levels - 1)


def round_reg(alpha, b, weight):  # block_recon.py:286-291
    return weight * (1 - ((soft_targets(alpha) - 0.5).abs() * 2).pow(b)).sum()


# ---- qdiff/quant_layer.py:406-437 -----------------------------------------------------------------
def quant_module_forward(x, weight, bias, kind, fwd_kwargs, act_q, w_q, split=0):
    """act_q / w_q: list of one (or, with split, two) dicts {delta, zero_point, n_levels[, keep][, alpha, soft]}.
    act_q None -> activation quantization off."""
    def qa(t, q):
        return uaq_forward(t, q["delta"], q["zero_point"], q["n_levels"], q.get("keep"))

    def qw(t, q):
        if q.get("alpha") is not None:
            return adaround_forward(t, q["alpha"], q["delta"], q["zero_point"], q["n_levels"], q.get("soft", False))
        return uaq_forward(t, q["delta"], q["zero_point"], q["n_levels"])

    if act_q is not None:
        if split:
            x = torch.cat([qa(x[:, :split], act_q[0]), qa(x[:, split:], act_q[1])], dim=1)
        else:
            x = qa(x, act_q[0])
    if w_q is not None:
        if split:
            weight = torch.cat([qw(weight[:, :split], w_q[0]), qw(weight[:, split:], w_q[1])], dim=1)
        else:
            weight = qw(weight, w_q[0])
    fn = {"conv2d": F.conv2d, "conv1d": F.conv1d, "linear": F.linear}[kind]
    return fn(x, weight, bias, **fwd_kwargs)


# ---- qdiff/quant_block.py:419-451 (QuantAttnBlock attention core, CIFAR) ------------------------------
def attn_core_ddim(q, k, v, quant):
    """q,k,v: [b,c,h,w] conv outputs.  quant: None or dict of four quantizer dicts q,k,v,w."""
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    if quant is not None:
        q = uaq_forward(q, **quant["q"])
        k = uaq_forward(k, **quant["k"])
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    w_ = w_.permute(0, 2, 1)
    if quant is not None:
        v = uaq_forward(v, **quant["v"])
        w_ = uaq_forward(w_, **quant["w"])
    return torch.bmm(v, w_).reshape(b, c, h, w)


# ---- qdiff/quant_block.py:119-165 + openaimodel.py:373-405 (LDM legacy attention) ---------------------
def attn_core_ldm(qkv, n_heads, quant):
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    q, k = q * scale, k * scale
    if quant is not None:
        q = uaq_forward(q, **quant["q"])
        k = uaq_forward(k, **quant["k"])
    weight = torch.einsum("bct,bcs->bts", q, k)
    weight = torch.softmax(weight.float(), dim=-1)
    if quant is not None:
        weight = uaq_forward(weight, **quant["w"])
        v = uaq_forward(v, **quant["v"])
    a = torch.einsum("bts,bcs->bct", weight, v)
    return a.reshape(bs, -1, length)


# ---- qdiff/quant_block.py:204-235 (cross_attn_forward core, after to_q/k/v, before to_out) ------------
def attn_core_cross(q, k, v, heads, scale, quant):
    b, n, hd = q.shape
    d = hd // heads

    def split_heads(t):
        return t.reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(t.shape[0] * heads, t.shape[1], d)

    q, k, v = split_heads(q), split_heads(k), split_heads(v)
    if quant is not None:
        q = uaq_forward(q, **quant["q"])
        k = uaq_forward(k, **quant["k"])
    sim = torch.einsum("bid,bjd->bij", q, k) * scale
    attn = sim.softmax(dim=-1)
    if quant is not None:
        attn = uaq_forward(attn, **quant["w"])
        v = uaq_forward(v, **quant["v"])
    out = torch.einsum("bij,bjd->bid", attn, v)
    return out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, heads * d)


# ---- qdiff/block_recon.py:305-323 -----------------------------------------------------------------
def linear_temp_decay(t, t_max, rel_start_decay, start_b, end_b):
    start_decay = rel_start_decay * t_max
    if t < start_decay:
        return start_b
    rel_t = (t - start_decay) / (t_max - start_decay)
    return end_b + (start_b - end_b) * max(0.0, 1 - rel_t)
