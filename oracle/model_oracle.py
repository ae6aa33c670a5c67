"""Model-level CPU oracle (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Wraps an FP UNet (the L0 structures in eda-dm_b200/unet_zoo, or the reference's own classes: only attribute
names are used) with plain-torch fake-quant layers that restate the reference's QuantModel
(qdiff/quant_model.py:12-95), QuantModule.forward (quant_layer.py:406-437), QuantAttnBlock.forward
(quant_block.py:419-451), QuantQKMatMul / QuantSMVMatMul (quant_block.py:119-165) and cross_attn_forward
(quant_block.py:204-235).  Quantizer naming follows the reference's module paths so golden (delta, zero_point)
tables recorded from the reference map 1:1.

Used by tests (checker), by smoke() (checker) and by bench.py's cpu_baseline / --impl reference legs (the timed
CPU baseline).  Never imported by the product.
"""
import math
from types import MethodType

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import qdiff_oracle as O


class OQuantizer:
    """State of one UniformAffineQuantizer (quant_layer.py:36-76)."""

    def __init__(self, n_bits=8, symmetric=False, channel_wise=False, scale_method='mse', leaf_param=False,
                 always_zero=False, prob=1.0):
        self.n_bits, self.sym, self.channel_wise, self.leaf_param = n_bits, symmetric, channel_wise, leaf_param
        self.delta = self.zero_point = None
        self.inited = False
        self.running = {} if leaf_param else None
        self.one_side = None

    @property
    def n_levels(self):
        return 2 ** self.n_bits

    def __call__(self, x):
        if not self.inited:  # quant_layer.py:247-264: search on every un-inited forward
            if self.one_side is None:
                self.one_side = "pos" if x.min() >= 0.0 else "neg" if x.max() <= 0.0 else "no"
            n_levels = self.n_levels
            best_min, best_max = O.search_1d(x, n_levels, self.channel_wise, self.one_side)
            if self.running is not None:
                if self.running.get("min") is None:
                    self.running["min"], self.running["max"] = best_min, best_max
                self.running["min"] = 0.1 * best_min + 0.9 * self.running["min"]
                self.running["max"] = 0.1 * best_max + 0.9 * self.running["max"]
                best_min, best_max = self.running["min"], self.running["max"]
            d, z = O.calculate_qparams(best_min, best_max, n_levels)
            if self.channel_wise:
                shape = [1] * x.dim()
                shape[0] = x.shape[0]
                d, z = d.reshape(shape), z.reshape(shape)
            self.delta, self.zero_point = d, z
        return O.uaq_forward(x, self.delta, self.zero_point, self.n_levels)

    def cheap_init(self, x):
        """max-abs range (NOT the reference's search) -- only for timing runs where the values of delta do not
        change the arithmetic being timed."""
        if self.channel_wise:
            m = torch.flatten(x, 1).abs().amax(1)
            shape = [1] * x.dim()
            shape[0] = x.shape[0]
            self.delta = (2 * m / (self.n_levels - 1)).clamp_min(1e-8).reshape(shape)
            self.zero_point = torch.full_like(self.delta, float(self.n_levels // 2))
        else:
            if x.min() >= 0:
                self.delta = (x.max() / (self.n_levels - 1)).clamp_min(1e-8)
                self.zero_point = torch.zeros(())
            else:
                self.delta = (2 * x.abs().max() / (self.n_levels - 1)).clamp_min(1e-8)
                self.zero_point = torch.tensor(float(self.n_levels // 2))
        self.inited = True


class OQuantLayer(nn.Module):
    """QuantModule restated (quant_layer.py:360-446)."""

    def __init__(self, org, wq_params, aq_params):
        super().__init__()
        self.org = org
        if isinstance(org, nn.Conv2d):
            self.kind, self.kw = "conv2d", dict(stride=org.stride, padding=org.padding, dilation=org.dilation, groups=org.groups)
        elif isinstance(org, nn.Conv1d):
            self.kind, self.kw = "conv1d", dict(stride=org.stride, padding=org.padding, dilation=org.dilation, groups=org.groups)
        else:
            self.kind, self.kw = "linear", {}
        self.wq_params, self.aq_params = wq_params, aq_params
        self.weight_quantizer, self.act_quantizer = OQuantizer(**wq_params), OQuantizer(**aq_params)
        self.weight_quantizer_0 = self.act_quantizer_0 = None
        self.use_weight_quant = self.use_act_quant = False
        self.disable_act_quant = False
        self.split = 0

    def forward(self, x, split=0):
        if split != 0 and self.split == 0:
            self.split = split
            self.weight_quantizer_0, self.act_quantizer_0 = OQuantizer(**self.wq_params), OQuantizer(**self.aq_params)
        w, b = self.org.weight, self.org.bias
        if self.use_act_quant and not self.disable_act_quant:
            if self.split:
                x = torch.cat([self.act_quantizer(x[:, :self.split]), self.act_quantizer_0(x[:, self.split:])], dim=1)
            else:
                x = self.act_quantizer(x)
        if self.use_weight_quant:
            if self.split:
                w = torch.cat([self.weight_quantizer(w[:, :self.split]), self.weight_quantizer_0(w[:, self.split:])], dim=1)
            else:
                w = self.weight_quantizer(w)
        fn = {"conv2d": F.conv2d, "conv1d": F.conv1d, "linear": F.linear}[self.kind]
        return fn(x, w, b, **self.kw)


def _attn_quantizers(aq_params, sm_abit, softmax_overrides):
    pw = dict(aq_params)
    pw["n_bits"] = sm_abit
    pw.update(softmax_overrides)
    return dict(q=OQuantizer(**aq_params), k=OQuantizer(**aq_params), v=OQuantizer(**aq_params), w=OQuantizer(**pw))


def _ddim_attn_forward(self, x):  # quant_block.py:419-451
    h_ = self.norm(x)
    q, k, v = self.q(h_), self.k(h_), self.v(h_)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    if self.o_use_act_quant:
        q, k = self.o_quant["q"](q), self.o_quant["k"](k)
    w_ = F.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=2)
    v = v.reshape(b, c, h * w)
    w_ = w_.permute(0, 2, 1)
    if self.o_use_act_quant:
        v, w_ = self.o_quant["v"](v), self.o_quant["w"](w_)
    return x + self.proj_out(torch.bmm(v, w_).reshape(b, c, h, w))


def _ldm_qk_forward(self, q, k):  # quant_block.py:128-139
    q, k = q * self.scale, k * self.scale
    if self.o_use_act_quant:
        q, k = self.o_quant["q"](q), self.o_quant["k"](k)
    return torch.einsum("bct,bcs->bts", q, k)


def _ldm_smv_forward(self, weight, v):  # quant_block.py:157-162
    if self.o_use_act_quant:
        weight, v = self.o_quant["w"](weight), self.o_quant["v"](v)
    return torch.einsum("bts,bcs->bct", weight, v)


def _cross_attn_forward(self, x, context=None, mask=None):  # quant_block.py:204-235
    h = self.heads
    q = self.to_q(x)
    context = x if context is None else context
    k, v = self.to_k(context), self.to_v(context)

    def sp(t):
        b, n, hd = t.shape
        return t.reshape(b, n, h, hd // h).permute(0, 2, 1, 3).reshape(b * h, n, hd // h)

    q, k, v = sp(q), sp(k), sp(v)
    if self.o_use_act_quant:
        q, k = self.o_quant["q"](q), self.o_quant["k"](k)
    attn = (torch.einsum("bid,bjd->bij", q, k) * self.scale).softmax(dim=-1)
    if self.o_use_act_quant:
        attn, v = self.o_quant["w"](attn), self.o_quant["v"](v)
    out = torch.einsum("bij,bjd->bid", attn, v)
    bh, n, d = out.shape
    out = out.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)
    return self.to_out(out)


class OracleQuantUNet(nn.Module):
    """QuantModel restated: same traversal order as quant_module_refactor / quant_block_refactor."""

    def __init__(self, model, weight_quant_params, act_quant_params, sm_abit=8):
        super().__init__()
        self.model = model
        self.layers, self.attn_sites = [], []   # (reference-style name, object)
        self._wrap_layers(model, "model", weight_quant_params, act_quant_params)
        self._wrap_attention(model, "model", act_quant_params, sm_abit)

    def _wrap_layers(self, module, prefix, wq, aq):
        for name, child in module.named_children():
            path = f"{prefix}.{name}"
            if isinstance(child, (nn.Conv2d, nn.Conv1d, nn.Linear)):
                layer = OQuantLayer(child, wq, aq)
                setattr(module, name, layer)
                self.layers.append((path, layer))
            else:
                self._wrap_layers(child, path, wq, aq)

    def _wrap_attention(self, module, prefix, aq, sm_abit):
        for name, child in module.named_children():
            path = f"{prefix}.{name}"
            cls = type(child).__name__
            if cls == "AttnBlock":
                child.o_quant = _attn_quantizers(aq, sm_abit, {})
                child.o_use_act_quant = False
                child.forward = MethodType(_ddim_attn_forward, child)
                self.attn_sites.append((path, child, "act_quantizer_"))
            elif cls == "QKMatMul" and aq.get("leaf_param"):
                child.o_quant = dict(q=OQuantizer(**aq), k=OQuantizer(**aq))
                child.o_use_act_quant = False
                child.forward = MethodType(_ldm_qk_forward, child)
                self.attn_sites.append((path, child, "act_quantizer_"))
            elif cls == "SMVMatMul" and aq.get("leaf_param"):
                pw = dict(aq)
                pw.update(n_bits=sm_abit, symmetric=False, always_zero=True)
                child.o_quant = dict(v=OQuantizer(**aq), w=OQuantizer(**pw))
                child.o_use_act_quant = False
                child.forward = MethodType(_ldm_smv_forward, child)
                self.attn_sites.append((path, child, "act_quantizer_"))
            elif cls == "BasicTransformerBlock":
                for an in ("attn1", "attn2"):
                    attn = getattr(child, an)
                    attn.o_quant = _attn_quantizers(aq, sm_abit, dict(always_zero=True))
                    attn.o_use_act_quant = False
                    attn.forward = MethodType(_cross_attn_forward, attn)
                    self.attn_sites.append((f"{path}.{an}", attn, "act_quantizer_"))
                self._wrap_attention(child, path, aq, sm_abit)
            else:
                self._wrap_attention(child, path, aq, sm_abit)

    # ---- reference API restated -------------------------------------------------------------------
    def set_quant_state(self, weight_quant=False, act_quant=False):
        for _, l in self.layers:
            l.use_weight_quant, l.use_act_quant = weight_quant, act_quant
        for _, site, _p in self.attn_sites:
            site.o_use_act_quant = act_quant

    def set_first_last_layer_to_8bit(self):  # quant_model.py:77-88
        w_list = [l.weight_quantizer for _, l in self.layers]
        a_list = self._act_quantizers_in_module_order()
        w_list[0].n_bits = 8
        w_list[-1].n_bits = 8
        a_list[-2].n_bits = 8

    def _act_quantizers_in_module_order(self):
        """Order of `named_modules()` over the reference's QuantModel restricted to leaf_param quantizers; the last two
        entries are what set_first_last_layer_to_8bit touches.  In every supported UNet the network's final modules are
        plain QuantModules (norm_out/conv_out or out.2), preceded by the previous QuantModule's act quantizer."""
        return [l.act_quantizer for _, l in self.layers]

    def disable_network_output_quantization(self):  # quant_model.py:90-95
        self.layers[-1][1].disable_act_quant = True

    def forward(self, x, timesteps=None, context=None):
        return self.model(x, timesteps, context)

    # ---- quantizer tables -----------------------------------------------------------------------------
    def named_quantizers(self):
        out = {}
        for path, l in self.layers:
            out[f"{path}.weight_quantizer"] = l.weight_quantizer
            out[f"{path}.act_quantizer"] = l.act_quantizer
            if l.weight_quantizer_0 is not None:
                out[f"{path}.weight_quantizer_0"] = l.weight_quantizer_0
                out[f"{path}.act_quantizer_0"] = l.act_quantizer_0
        for path, site, pre in self.attn_sites:
            for k, q in site.o_quant.items():
                out[f"{path}.{pre}{k}"] = q
        return out

    def load_qparams(self, table):
        """table: {name: (delta, zero_point, n_bits)}; split twins are created on demand."""
        for path, l in self.layers:
            if f"{path}.weight_quantizer_0" in table and l.weight_quantizer_0 is None:
                l.weight_quantizer_0, l.act_quantizer_0 = OQuantizer(**l.wq_params), OQuantizer(**l.aq_params)
        named = self.named_quantizers()
        dev = next(self.model.parameters()).device
        for name, (d, z, bits) in table.items():
            q = named[name]
            q.delta, q.zero_point = torch.as_tensor(d).to(dev), torch.as_tensor(z).to(dev)
            q.n_bits, q.inited = int(bits), True

    def set_inited(self, flag, weights=True, acts=True):
        for name, q in self.named_quantizers().items():
            is_w = ".weight_quantizer" in name
            if (is_w and weights) or (not is_w and acts):
                q.inited = flag

    def cheap_calibrate(self, x, t, context=None):
        """One FP pass that records every quantizer input and sets max-abs ranges (timing runs only)."""
        hooks = []
        for _, l in self.layers:
            l.weight_quantizer.cheap_init(l.org.weight.detach())

            def pre(mod, args, l=l):
                xin = args[0]
                if l.split or (len(args) > 1 and args[1]):
                    s = l.split or args[1]
                    if l.act_quantizer_0 is None:
                        l.split = s
                        l.weight_quantizer_0, l.act_quantizer_0 = OQuantizer(**l.wq_params), OQuantizer(**l.aq_params)
                        l.weight_quantizer.cheap_init(l.org.weight.detach()[:, :s])
                        l.weight_quantizer_0.cheap_init(l.org.weight.detach()[:, s:])
                    l.act_quantizer.cheap_init(xin[:, :s])
                    l.act_quantizer_0.cheap_init(xin[:, s:])
                else:
                    l.act_quantizer.cheap_init(xin)
            hooks.append(l.register_forward_pre_hook(pre))
        self.set_quant_state(False, False)
        with torch.no_grad():
            self(x, t, context)
        for h in hooks:
            h.remove()
        for _, site, _p in self.attn_sites:
            for k, q in site.o_quant.items():
                # attention operands: q/k/v ~ activations of O(1); softmax probabilities in [0, 1]
                if k == "w":
                    q.delta, q.zero_point, q.inited = torch.tensor(1.0 / (q.n_levels - 1)), torch.zeros(()), True
                else:
                    q.delta, q.zero_point, q.inited = torch.tensor(8.0 / (q.n_levels - 1)), torch.tensor(float(q.n_levels // 2)), True
