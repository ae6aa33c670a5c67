"""Latent-diffusion UNet family (LSUN-Church / LSUN-Bedroom / ImageNet / Stable-Diffusion shapes).

FP structure only -- the L0 layer under the quantized path -- written from scratch with the public
ADM / latent-diffusion checkpoint naming (`time_embed`, `input_blocks.N.M`, `middle_block`,
`output_blocks`, `out`; `in_layers/emb_layers/out_layers/skip_connection`; `attn1/attn2/ff`), which is the
layout ldm/modules/diffusionmodules/openaimodel.py:447-783 and ldm/modules/attention.py of the reference
load, so reference state_dicts map 1:1.  The two attention matmuls are separate modules (QKMatMul /
SMVMatMul) because that is the seam the quantized path swaps (reference openaimodel.py:350-371).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([args.cos(), args.sin()], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def _zero(m):
    for p in m.parameters():
        p.detach().zero_()
    return m


class TimestepBlock(nn.Module):
    """Marker: forward(x, emb, split=0)."""


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if use_conv:
            self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        if self.use_conv and hasattr(self.conv, "forward_upsample2x"):     # QuantModule: upsample the u8 codes, not the fp32 tensor
            return self.conv.forward_upsample2x(x)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.conv(x) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        self.op = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding) if use_conv else nn.AvgPool2d(2, 2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 use_checkpoint=False, up=False, down=False):
        super().__init__()
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.updown = up or down
        if up:
            self.h_upd, self.x_upd = Upsample(channels, False), Upsample(channels, False)
        elif down:
            self.h_upd, self.x_upd = Downsample(channels, False), Downsample(channels, False)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        _zero(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb, split=0):
        return resblock_forward(self, x, emb, split)


def resblock_forward(blk, x, emb, split=0):
    """Shared by ResBlock and its quantized wrapper (same dataflow as reference quant_block.py:86-116)."""
    if blk.updown:
        h = blk.in_layers[:-1](x)
        h, x = blk.h_upd(h), blk.x_upd(x)
        h = blk.in_layers[-1](h)
    else:
        h = blk.in_layers(x)
    emb_out = blk.emb_layers(emb).type(h.dtype)
    while emb_out.dim() < h.dim():
        emb_out = emb_out[..., None]
    if blk.use_scale_shift_norm:
        scale, shift = torch.chunk(emb_out, 2, dim=1)
        h = blk.out_layers[0](h) * (1 + scale) + shift
        h = blk.out_layers[1:](h)
    else:
        h = blk.out_layers(h + emb_out)
    if split and not isinstance(blk.skip_connection, nn.Identity):
        return blk.skip_connection(x, split=split) + h
    return blk.skip_connection(x) + h


class QKMatMul(nn.Module):
    def __init__(self):
        super().__init__()
        self.scale = None

    def forward(self, q, k):
        return torch.einsum("bct,bcs->bts", q * self.scale, k * self.scale)


class SMVMatMul(nn.Module):
    def forward(self, weight, v):
        return torch.einsum("bts,bcs->bct", weight, v)


class QKVAttentionLegacy(nn.Module):
    def __init__(self, n_heads):
        super().__init__()
        self.n_heads = n_heads
        self.qkv_matmul = QKMatMul()
        self.smv_matmul = SMVMatMul()

    def forward(self, qkv):
        bs, width, length = qkv.shape
        ch = width // (3 * self.n_heads)
        q, k, v = qkv.reshape(bs * self.n_heads, ch * 3, length).split(ch, dim=1)
        self.qkv_matmul.scale = 1 / math.sqrt(math.sqrt(ch))
        weight = self.qkv_matmul(q, k)
        weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
        return self.smv_matmul(weight, v).reshape(bs, self.n_heads * ch, length)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads if num_head_channels == -1 else channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.attention = QKVAttentionLegacy(self.num_heads)
        self.proj_out = _zero(nn.Conv1d(channels, channels, 1))

    def forward(self, x):
        b, c, *spatial = x.shape
        x = x.flatten(2)
        # quantized modules (qdiff.QuantModule) take GroupNorm into their activation producer and the residual into
        # their epilogue; plain nn.Conv1d runs module by module
        if hasattr(self.qkv, 'forward_prenorm'):
            qkv = self.qkv.forward_prenorm(x, self.norm, silu=False)
        else:
            qkv = self.qkv(self.norm(x))
        h = self.attention(qkv)
        if hasattr(self.proj_out, 'forward_prenorm') and not self.proj_out._forward_hooks:
            return self.proj_out(h, residual=x).reshape(b, c, *spatial)
        return (x + self.proj_out(h)).reshape(b, c, *spatial)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, glu=True, dropout=0.0):
        super().__init__()
        inner = int(dim * mult)
        first = GEGLU(dim, inner) if glu else nn.Sequential(nn.Linear(dim, inner), nn.GELU())
        self.net = nn.Sequential(first, nn.Dropout(dropout), nn.Linear(inner, dim))

    def forward(self, x):
        return self.net(x)


def _heads_split(t, h):
    b, n, hd = t.shape
    return t.reshape(b, n, h, hd // h).permute(0, 2, 1, 3).reshape(b * h, n, hd // h)


def _heads_merge(t, h):
    bh, n, d = t.shape
    return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim or query_dim
        self.scale, self.heads = dim_head ** -0.5, heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))

    def forward(self, x, context=None, mask=None):
        context = x if context is None else context
        q, k, v = (_heads_split(t, self.heads) for t in (self.to_q(x), self.to_k(context), self.to_v(context)))
        attn = (torch.einsum("bid,bjd->bij", q, k) * self.scale).softmax(dim=-1)
        return self.to_out(_heads_merge(torch.einsum("bij,bjd->bid", attn, v), self.heads))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=True):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.checkpoint = checkpoint

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None):
        super().__init__()
        self.in_channels = in_channels
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, dropout=dropout, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = _zero(nn.Conv2d(inner, in_channels, 1))

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        # quantized 1x1 convs (qdiff.QuantModule) emit / consume the token layout directly and take the GroupNorm and the
        # residual with them; plain nn.Conv2d runs module by module
        if hasattr(self.proj_in, 'forward_prenorm'):
            y = self.proj_in.forward_prenorm(x, self.norm, silu=False, tokens_out=True)
        else:
            y = self.proj_in(self.norm(x)).flatten(2).permute(0, 2, 1)
        for blk in self.transformer_blocks:
            y = blk(y, context)
        if hasattr(self.proj_out, 'forward_from_tokens') and not self.proj_out._forward_hooks:
            return self.proj_out.forward_from_tokens(y, (h, w), residual=x)
        y = y.permute(0, 2, 1).reshape(b, y.shape[2], h, w)
        return self.proj_out(y) + x


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None, split=0):
        for layer in self:
            if isinstance(layer, TimestepBlock) or getattr(layer, "takes_emb", False):
                x = layer(x, emb, split=split)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, num_classes=None, use_checkpoint=False,
                 num_heads=-1, num_head_channels=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_spatial_transformer=False, transformer_depth=1, context_dim=None, legacy=True, **_ignored):
        super().__init__()
        self.image_size, self.in_channels, self.model_channels, self.out_channels = image_size, in_channels, model_channels, out_channels
        self.num_classes = num_classes
        self.split_shortcut = False
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, ted)

        def res(cin, cout, **kw):
            return ResBlock(cin, ted, dropout, out_channels=cout, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, **kw)

        def attn(ch):
            if num_head_channels == -1:
                heads, dim_head = num_heads, ch // num_heads
            else:
                heads, dim_head = ch // num_head_channels, num_head_channels
            if legacy:
                dim_head = ch // heads if use_spatial_transformer else num_head_channels
            if use_spatial_transformer:
                return SpatialTransformer(ch, heads, dim_head, depth=transformer_depth, context_dim=context_dim)
            return AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=heads, num_head_channels=dim_head)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(
                    res(ch, ch, down=True) if resblock_updown else Downsample(ch, conv_resample, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch), res(ch, ch))
        self.output_blocks = nn.ModuleList()
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, up=True) if resblock_updown else Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), _zero(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    # The forward pass as a list of stages over an explicit state {h, hs, emb, context}: `forward` runs them all; the calibration
    # cache builder (qdiff/data_utils.py StagedCache) keeps the state of every calibration batch at the frontier of the units
    # already reconstructed and only ever runs each stage once per path, instead of re-running the whole prefix for every unit.
    def stage_modules(self):
        return ([[self.time_embed] + ([self.label_emb] if self.num_classes is not None else [])] + [[m] for m in self.input_blocks] +
                [[self.middle_block]] + [[m] for m in self.output_blocks] + [[self.out]])

    def stage_begin(self, x, timesteps=None, context=None, y=None, **kwargs):
        return {"x": x, "timesteps": timesteps, "context": context, "y": y, "h": x, "hs": [], "emb": None}

    def run_stage(self, k, st):
        st = dict(st)
        n_in = len(self.input_blocks)
        if k == 0:
            emb = self.time_embed(timestep_embedding(st["timesteps"], self.model_channels))
            if self.num_classes is not None:
                emb = emb + self.label_emb(st["y"])
            st["emb"] = emb
        elif k <= n_in:
            st["h"] = self.input_blocks[k - 1](st["h"], st["emb"], st["context"])
            st["hs"] = st["hs"] + [st["h"]]
        elif k == n_in + 1:
            st["h"] = self.middle_block(st["h"], st["emb"], st["context"])
        elif k <= n_in + 1 + len(self.output_blocks):
            module = self.output_blocks[k - n_in - 2]
            h = st["h"]
            split = h.shape[1] if self.split_shortcut else 0
            # a quantized ResBlock on the integer path reads the two sources of the skip concatenation in place (lazy_cat)
            lazy = getattr(module[0], "lazy_cat", None) if not (module._forward_hooks or module._forward_pre_hooks) else None
            inp = lazy(h, st["hs"][-1], split) if lazy is not None else None
            if inp is None:
                inp = torch.cat([h, st["hs"][-1]], dim=1)
            st["h"] = module(inp, st["emb"], st["context"], split=split)
            st["hs"] = st["hs"][:-1]
        else:
            h = st["h"]
            if hasattr(self.out[-1], 'forward_prenorm') and len(self.out) == 3 and isinstance(self.out[1], nn.SiLU):
                st["h"] = self.out[2].forward_prenorm(h, self.out[0])      # GroupNorm + SiLU ride on the quantized module's producer
            else:
                st["h"] = self.out(h)
        return st

    def forward(self, x, timesteps=None, context=None, y=None, **kwargs):
        st = self.stage_begin(x, timesteps, context, y)
        for k in range(len(self.input_blocks) + len(self.output_blocks) + 3):
            st = self.run_stage(k, st)
        return st["h"]


def reinit_zero_modules(model: nn.Module, std: float = 0.02, seed: int = 0):
    """The public init zeroes some convs (`zero_module`); synthetic benchmarks re-draw them so that no GEMM
    is trivially zero."""
    g = torch.Generator().manual_seed(seed)
    for p in model.parameters():
        if p.dim() > 1 and float(p.detach().abs().max()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=g) * std)
    return model


# unet_config.params of the reference's model configs (file:line in each docstring)
def lsun_church_unet():
    """models/ldm/lsun_churches256/config.yaml:32-53 (LDM-8, latent 32x32x4, 295.0 M params)."""
    return UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=192, attention_resolutions=[1, 2, 4, 8],
                     num_res_blocks=2, channel_mult=[1, 2, 2, 4, 4], num_heads=8, use_scale_shift_norm=True, resblock_updown=True)


def lsun_bedroom_unet():
    """models/ldm/lsun_beds256/config.yaml:17-34 (LDM-4, latent 64x64x3, 274.1 M params)."""
    return UNetModel(image_size=64, in_channels=3, out_channels=3, model_channels=224, attention_resolutions=[8, 4, 2],
                     num_res_blocks=2, channel_mult=[1, 2, 3, 4], num_head_channels=32)


def imagenet_unet():
    """configs/latent-diffusion/cin256-v2.yaml:19-39 (LDM-4 class-conditional, 400.9 M params)."""
    return UNetModel(image_size=64, in_channels=3, out_channels=3, model_channels=192, attention_resolutions=[8, 4, 2],
                     num_res_blocks=2, channel_mult=[1, 2, 3, 5], num_heads=1, use_spatial_transformer=True,
                     transformer_depth=1, context_dim=512)


def stable_diffusion_unet():
    """configs/stable-diffusion/v1-inference.yaml:29-44 (SD v1.4, 859.5 M params)."""
    return UNetModel(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                     num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                     transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False)
