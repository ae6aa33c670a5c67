"""DDPM/DDIM pixel-space UNet (the CIFAR-10 architecture of BASELINE config 0).

FP structure only (GroupNorm, swish, residual adds, up/down-sampling): this is the L0 layer the
quantized path wraps, written from scratch so that synthetic-weight benchmarks and tests do not need
the reference tree.  Parameter names and shapes follow the public DDPM checkpoint layout (the same one
ddim/models/diffusion.py:199-308 of the reference loads), so a reference state_dict loads unchanged.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * (-math.log(10000.0) / (half - 1)))
    args = t.float()[:, None] * freqs[None, :]
    emb = torch.cat([args.sin(), args.cos()], dim=1)
    if dim % 2:
        emb = F.pad(emb, (0, 1))
    return emb


def nonlinearity(x):
    return x * torch.sigmoid(x)


def _gn(ch):
    return nn.GroupNorm(32, ch, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, ch, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(ch, ch, 3, 1, 1)

    def forward(self, x):
        if self.with_conv and hasattr(self.conv, "forward_upsample2x"):    # QuantModule: upsample the u8 codes, not the fp32 tensor
            return self.conv.forward_upsample2x(x)
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, ch, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(ch, ch, 3, 2, 0)

    def forward(self, x):
        if not self.with_conv:
            return F.avg_pool2d(x, 2, 2)
        return self.conv(F.pad(x, (0, 1, 0, 1)))  # asymmetric pad, stride-2 valid conv


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = _gn(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = _gn(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, x, temb=None, split=0):
        h = self.conv1(nonlinearity(self.norm1(x)))
        h = h + self.temb_proj(nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(nonlinearity(self.norm2(h))))
        if self.in_channels != self.out_channels:
            if self.use_conv_shortcut:
                x = self.conv_shortcut(x)
            else:
                x = self.nin_shortcut(x, split) if split else self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = _gn(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)

    def forward(self, x):
        h = self.norm(x)
        q, k, v = self.q(h), self.k(h), self.v(h)
        b, c, hh, ww = q.shape
        q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
        k = k.reshape(b, c, hh * ww)
        w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
        v = v.reshape(b, c, hh * ww)
        h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class DDPMUNet(nn.Module):
    def __init__(self, ch=128, out_ch=3, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,), dropout=0.1,
                 in_channels=3, resolution=32, resamp_with_conv=True):
        super().__init__()
        # `config.split_shortcut` is the switch scripts/sample_diffusion_ddim.py:286 flips on the wrapped model
        self.config = SimpleNamespace(split_shortcut=False, change_block_recon=False)
        self.ch, self.temb_ch = ch, ch * 4
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        ch_mult = tuple(ch_mult)

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)

        res = resolution
        in_mult = (1,) + ch_mult
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(self.num_resolutions):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_mult[lvl], ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                level.block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if res in attn_resolutions:
                    level.attn.append(AttnBlock(block_in))
            if lvl != self.num_resolutions - 1:
                level.downsample = Downsample(block_in, resamp_with_conv)
                res //= 2
            self.down.append(level)

        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=self.temb_ch, dropout=dropout)

        self.up = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            level = nn.Module()
            level.block, level.attn = nn.ModuleList(), nn.ModuleList()
            block_out, skip_in = ch * ch_mult[lvl], ch * ch_mult[lvl]
            for i in range(num_res_blocks + 1):
                if i == num_res_blocks:
                    skip_in = ch * in_mult[lvl]
                level.block.append(ResnetBlock(in_channels=block_in + skip_in, out_channels=block_out, temb_channels=self.temb_ch, dropout=dropout))
                block_in = block_out
                if res in attn_resolutions:
                    level.attn.append(AttnBlock(block_in))
            if lvl != 0:
                level.upsample = Upsample(block_in, resamp_with_conv)
                res *= 2
            self.up.insert(0, level)

        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)

    # stages over an explicit state (see unet_zoo/ldm_unet.py UNetModel.stage_modules): embedding, conv_in, every
    # (ResnetBlock [+ AttnBlock]) pair and resampling layer in execution order, the middle, the output head
    def _plan(self):
        plan = [("temb", None, None), ("conv_in", None, None)]
        for lvl in range(self.num_resolutions):
            for i in range(self.num_res_blocks):
                plan.append(("down", lvl, i))
            if lvl != self.num_resolutions - 1:
                plan.append(("downsample", lvl, None))
        plan.append(("mid", None, None))
        for lvl in reversed(range(self.num_resolutions)):
            for i in range(self.num_res_blocks + 1):
                plan.append(("up", lvl, i))
            if lvl != 0:
                plan.append(("upsample", lvl, None))
        plan.append(("out", None, None))
        return plan

    def stage_modules(self):
        mods = []
        for kind, lvl, i in self._plan():
            if kind == "temb":
                mods.append(list(self.temb.dense))
            elif kind == "conv_in":
                mods.append([self.conv_in])
            elif kind in ("down", "up"):
                level = (self.down if kind == "down" else self.up)[lvl]
                mods.append([level.block[i]] + ([level.attn[i]] if len(level.attn) > 0 else []))
            elif kind == "downsample":
                mods.append([self.down[lvl].downsample])
            elif kind == "upsample":
                mods.append([self.up[lvl].upsample])
            elif kind == "mid":
                mods.append([self.mid.block_1, self.mid.attn_1, self.mid.block_2])
            else:
                mods.append([self.norm_out, self.conv_out])
        return mods

    def stage_begin(self, x, t=None, context=None):
        return {"x": x, "t": t, "h": None, "hs": [], "temb": None}

    def run_stage(self, k, st):
        st = dict(st)
        kind, lvl, i = self._plan()[k]
        if kind == "temb":
            st["temb"] = self.temb.dense[1](nonlinearity(self.temb.dense[0](sinusoidal_embedding(st["t"], self.ch))))
        elif kind == "conv_in":
            st["hs"] = [self.conv_in(st["x"])]
        elif kind == "down":
            level = self.down[lvl]
            h = level.block[i](st["hs"][-1], st["temb"])
            if len(level.attn) > 0:
                h = level.attn[i](h)
            st["hs"] = st["hs"] + [h]
        elif kind == "downsample":
            st["hs"] = st["hs"] + [self.down[lvl].downsample(st["hs"][-1])]
        elif kind == "mid":
            st["h"] = self.mid.block_2(self.mid.attn_1(self.mid.block_1(st["hs"][-1], st["temb"])), st["temb"])
        elif kind == "up":
            level, h = self.up[lvl], st["h"]
            # split shortcut: the concat boundary is handed to the 1x1 shortcut conv so the two halves get
            # their own quantizers (reference ddim/models/diffusion.py:357-368)
            split = h.size(1) if self.config.split_shortcut else 0
            cat = torch.cat([h, st["hs"][-1]], dim=1)
            h = level.block[i](cat, st["temb"], split=split) if split else level.block[i](cat, st["temb"])
            if len(level.attn) > 0:
                h = level.attn[i](h)
            st["h"], st["hs"] = h, st["hs"][:-1]
        elif kind == "upsample":
            st["h"] = self.up[lvl].upsample(st["h"])
        else:
            h = st["h"]
            if hasattr(self.conv_out, 'forward_prenorm'):
                st["h"] = self.conv_out.forward_prenorm(h, self.norm_out, act_fn=nonlinearity)
            else:
                st["h"] = self.conv_out(nonlinearity(self.norm_out(h)))
        return st

    def forward(self, x, t=None, context=None):
        st = self.stage_begin(x, t, context)
        for k in range(len(self._plan())):
            st = self.run_stage(k, st)
        return st["h"]


def cifar10_unet(**overrides):
    """configs/cifar10.yml of the reference: 35.7 M parameters."""
    kw = dict(ch=128, out_ch=3, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,), dropout=0.1,
              in_channels=3, resolution=32, resamp_with_conv=True)
    kw.update(overrides)
    return DDPMUNet(**kw)
