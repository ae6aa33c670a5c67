"""CPU restatement of ONE block_reconstruction iteration (TEST INFRASTRUCTURE / timed CPU baseline -- see oracle/__init__.py).

Follows the reference loop body qdiff/block_recon.py:133-206 on the DDIM ResnetBlock `down.0.block.0` of the CIFAR-10 UNet
(timed section (ii)): input mixing with QDrop probability 0.5 (:141-145), quantized forward (:152-155), FP
forward and second quantized forward with per-layer taps (:161-171), FBR loss over all but the last layer (:188-191), block loss
(:193-195, LossFunction :259-302 with round_loss='none'), backward (:197), Adam steps on the AdaRound alphas and the activation
step sizes (:199-206).  Quantizer arithmetic comes from oracle/qdiff_oracle.py (UniformAffineQuantizer.forward
quant_layer.py:267-274, AdaRoundQuantizer.forward adaptive_rounding.py:49-59).

Parity status: the arithmetic pieces are pinned by tests/test_oracle_golden.py; the assembled iteration is a TIMING baseline only
(the product's loss trajectories are pinned directly against traces recorded from the reference: tests/test_gpu_model.py).
Also times section (iii): the activation range search of set_act_quantize_params (quant_layer.py:150-213 via O.search_1d).
"""
import statistics
import time

import torch
import torch.nn.functional as F

from . import qdiff_oracle as O


class _Layer:
    """one QuantModule of the unit: soft AdaRound weights + trainable activation step size"""

    def __init__(self, weight, bias, kind, kw, gen):
        self.weight, self.bias, self.kind, self.kw = weight, bias, kind, kw
        flat = weight.flatten(1)
        self.dw = (2 * flat.abs().amax(1) / 15).clamp_min(1e-8).reshape(-1, *([1] * (weight.dim() - 1)))
        self.zw = torch.full_like(self.dw, 8.0)
        self.alpha = O.adaround_init_alpha(weight, self.dw).requires_grad_(True)
        self.da = torch.tensor(0.05, requires_grad=True)
        self.za = torch.tensor(128.0)
        self.gen = gen

    def __call__(self, x, quant, taps):
        if quant:
            keep = torch.rand(x.shape, generator=self.gen) < 0.5            # QDrop, quant_layer.py:271-272
            x = O.uaq_forward(x, self.da, self.za, 256, keep=keep)
            w = O.adaround_forward(self.weight, self.alpha, self.dw, self.zw, 16, soft=True)
        else:
            w = self.weight
        out = (F.conv2d if self.kind == "conv" else F.linear)(x, w, self.bias, **self.kw)
        taps.append(out)
        return out


def _block(layers, norms, x, temb, quant, taps):
    conv1, temb_proj, conv2 = layers
    h = conv1(F.silu(norms[0](x)), quant, taps)
    h = h + temb_proj(F.silu(temb), quant, taps)[:, :, None, None]
    h = conv2(F.silu(norms[1](h)), quant, taps)
    return x + h


def cpu_recon_baseline(threads, batch=32, ch=128, res=32, temb_ch=512, iters=5, search_elems=1 << 20):
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    rnd = lambda *s: torch.randn(*s, generator=g)
    layers = [_Layer(rnd(ch, ch, 3, 3) * 0.03, rnd(ch) * 0.01, "conv", dict(padding=1), g),
              _Layer(rnd(ch, temb_ch) * 0.03, rnd(ch) * 0.01, "linear", {}, g),
              _Layer(rnd(ch, ch, 3, 3) * 0.03, rnd(ch) * 0.01, "conv", dict(padding=1), g)]
    norms = [torch.nn.GroupNorm(32, ch, eps=1e-6), torch.nn.GroupNorm(32, ch, eps=1e-6)]
    w_opt = torch.optim.Adam([l.alpha for l in layers], lr=1e-2)
    a_opt = torch.optim.Adam([l.da for l in layers], lr=4e-4)
    inp_q, inp_fp, temb, target = rnd(batch, ch, res, res), rnd(batch, ch, res, res), rnd(batch, temb_ch), rnd(batch, ch, res, res)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        cur = torch.where(torch.rand(inp_q.shape, generator=g) < 0.5, inp_q, inp_fp)          # block_recon.py:141-145
        w_opt.zero_grad(); a_opt.zero_grad()
        out_quant = _block(layers, norms, cur, temb, True, [])                               # :152-155
        r_taps, q_taps = [], []
        with torch.no_grad():
            _block(layers, norms, inp_fp, temb, False, r_taps)                               # :161-165
        _block(layers, norms, cur, temb, True, q_taps)                                       # :167-171
        m_loss = sum(O.lp_loss(q_taps[j], r_taps[j], p=2) for j in range(len(r_taps) - 1))  # :188-191
        loss = O.lp_loss(out_quant, target, p=2.0) + 0.8 * m_loss                            # :193-195
        loss.backward()
        w_opt.step(); a_opt.step()
        if it > 0:
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    # (iii) one per-tensor activation range search (100 candidates, quant_layer.py:150-213) on a 1 Mi-element tensor
    x = rnd(search_elems)
    t0 = time.perf_counter()
    O.search_1d(x, 256, False, "no")
    t_search = time.perf_counter() - t0
    return {"block_recon_iters_per_s": 1.0 / med, "ms_per_iter": med * 1e3, "cores": threads, "kind": "port",
            "sample": f"{iters} iterations of the reference loop body on a CIFAR DDIM ResnetBlock ({ch} ch, {res}x{res}, batch {batch}), "
                      f"median; oracle/recon_oracle.py",
            "act_range_search_s_per_Mi_elements": t_search,
            "loss_finite": bool(torch.isfinite(loss))}
