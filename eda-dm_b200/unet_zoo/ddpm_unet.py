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

    def forward(self, x, t=None, context=None):
        temb = self.temb.dense[1](nonlinearity(self.temb.dense[0](sinusoidal_embedding(t, self.ch))))
        hs = [self.conv_in(x)]
        for lvl in range(self.num_resolutions):
            level = self.down[lvl]
            for i in range(self.num_res_blocks):
                h = level.block[i](hs[-1], temb)
                if len(level.attn) > 0:
                    h = level.attn[i](h)
                hs.append(h)
            if lvl != self.num_resolutions - 1:
                hs.append(level.downsample(hs[-1]))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], temb)), temb)
        for lvl in reversed(range(self.num_resolutions)):
            level = self.up[lvl]
            for i in range(self.num_res_blocks + 1):
                # split shortcut: the concat boundary is handed to the 1x1 shortcut conv so the two halves get
                # their own quantizers (reference ddim/models/diffusion.py:357-368)
                split = h.size(1) if self.config.split_shortcut else 0
                h = level.block[i](torch.cat([h, hs.pop()], dim=1), temb, split=split) if split else \
                    level.block[i](torch.cat([h, hs.pop()], dim=1), temb)
                if len(level.attn) > 0:
                    h = level.attn[i](h)
            if lvl != 0:
                h = level.upsample(h)
        if hasattr(self.conv_out, 'forward_prenorm'):
            return self.conv_out.forward_prenorm(h, self.norm_out, act_fn=nonlinearity)
        return self.conv_out(nonlinearity(self.norm_out(h)))


def cifar10_unet(**overrides):
    """configs/cifar10.yml of the reference: 35.7 M parameters."""
    kw = dict(ch=128, out_ch=3, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,), dropout=0.1,
              in_channels=3, resolution=32, resamp_with_conv=True)
    kw.update(overrides)
    return DDPMUNet(**kw)
