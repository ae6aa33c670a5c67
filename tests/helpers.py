"""Shared helpers: load golden fixtures into the from-scratch UNet structures."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

WQ = {'n_bits': 4, 'symmetric': True, 'channel_wise': True, 'scale_method': 'mse'}
AQ = {'n_bits': 8, 'symmetric': True, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True, 'prob': 1.0}

LDM_KW = {
    "ldm_tiny.npz": dict(image_size=8, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1, 2],
                         num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_scale_shift_norm=True, resblock_updown=True),
    "ldm_tiny_b.npz": dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[2],
                           num_res_blocks=1, channel_mult=[1, 2], num_head_channels=16),
    "ldm_xattn_tiny.npz": dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[1, 2],
                               num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True,
                               transformer_depth=1, context_dim=24),
}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def state_dict(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


def qtable(g):
    names = [str(n) for n in g["q.names"]]
    return {n: (torch.from_numpy(g[f"q.{i}_delta"]), torch.from_numpy(g[f"q.{i}_zp"]), int(g[f"q.{i}_bits"]))
            for i, n in enumerate(names)}


def ddim_tiny_model():
    from unet_zoo.ddpm_unet import DDPMUNet
    return DDPMUNet(ch=32, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), dropout=0.0, in_channels=3,
                    resolution=16, resamp_with_conv=True).eval()


def ldm_model(name):
    from unet_zoo.ldm_unet import UNetModel
    return UNetModel(**LDM_KW[name]).eval()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def install_qparams(qnn, table):
    """Copy a golden (delta, zero_point, n_bits) table into a product QuantModel (names follow the reference).
    Split twins must already exist (run one forward with split_shortcut on first)."""
    from qdiff.quant_layer import UniformAffineQuantizer
    dev = next(qnn.parameters()).device
    named = dict(qnn.named_modules())
    for name, (d, z, bits) in table.items():
        q = named[name]
        assert isinstance(q, UniformAffineQuantizer), name
        q.bitwidth_refactor(bits)
        q.zero_point = z.to(dev)
        q.delta = torch.nn.Parameter(d.to(dev)) if q.leaf_param else d.to(dev)
        q.inited = True


def qtable_for_oracle(g, om):
    """golden table restricted to the names the oracle model knows (drops the unused BaseQuantBlock.act_quantizer
    entries, which never get a delta anyway)."""
    named = om.named_quantizers()
    return {k: v for k, v in qtable(g).items() if k in named}
