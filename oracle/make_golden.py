"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

TEST INFRASTRUCTURE.  Run once in the build container (`python oracle/make_golden.py`); the outputs are
committed.  The reference has no tests or golden vectors of its own, so these files are
what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path (tests -m gpu).
"""
import argparse
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
warnings.filterwarnings("ignore")

import torch  # noqa: E402
from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params  # noqa: E402  (the REFERENCE's qdiff)
from qdiff.quant_layer import UniformAffineQuantizer, QuantModule  # noqa: E402
from qdiff.adaptive_rounding import AdaRoundQuantizer  # noqa: E402
from qdiff.quant_block import QuantQKMatMul, QuantSMVMatMul, QuantBasicTransformerBlock, QuantAttnBlock  # noqa: E402
import qdiff.block_recon as ref_block_recon  # noqa: E402
import qdiff.layer_recon as ref_layer_recon  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
WQ = {'n_bits': 4, 'symmetric': True, 'channel_wise': True, 'scale_method': 'mse'}
AQ = {'n_bits': 8, 'symmetric': True, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True, 'prob': 1.0}


def npy(t):
    return t.detach().cpu().numpy().copy()


def ns(d):
    n = argparse.Namespace()
    for k, v in d.items():
        setattr(n, k, ns(v) if isinstance(v, dict) else v)
    return n


def quantizer_table(qnn):
    """name -> (delta, zero_point, n_bits) for every inited quantizer of the reference model"""
    tab = {}
    for name, m in qnn.named_modules():
        if isinstance(m, (UniformAffineQuantizer, AdaRoundQuantizer)) and m.delta is not None:
            tab[name] = (npy(m.delta), npy(m.zero_point), int(m.n_bits))
    return tab


def pack_table(prefix, tab, out):
    out[prefix + "names"] = np.array(sorted(tab))
    for i, k in enumerate(sorted(tab)):
        out[f"{prefix}{i}_delta"], out[f"{prefix}{i}_zp"], out[f"{prefix}{i}_bits"] = tab[k][0], tab[k][1], np.int64(tab[k][2])


def unit_vectors():
    g = torch.Generator().manual_seed(11)
    out = {}
    # per-tensor activation quantizer with EMA over two batches (leaf_param), two-sided
    q = UniformAffineQuantizer(**AQ)
    xs = [torch.randn(8, 16, 8, 8, generator=g) * 1.3, torch.randn(8, 16, 8, 8, generator=g) * 0.9]
    ys = [q(x) for x in xs]
    q.set_inited(True)
    out.update(act_x0=npy(xs[0]), act_x1=npy(xs[1]), act_y0=npy(ys[0]), act_y1=npy(ys[1]), act_delta=npy(q.delta),
               act_zp=npy(q.zero_point))
    # one-sided positive (softmax-like), sm quantizer settings of QuantSMVMatMul
    pw = dict(AQ); pw.update(symmetric=False, always_zero=True)
    q = UniformAffineQuantizer(**pw)
    xp = torch.softmax(torch.randn(4, 32, 32, generator=g), -1)
    yp = q(xp)
    out.update(pos_x=npy(xp), pos_y=npy(yp), pos_delta=npy(q.delta), pos_zp=npy(q.zero_point))
    # channel-wise weights, 4 and 8 bit
    w = torch.randn(24, 16, 3, 3, generator=g) * 0.08
    for bits in (4, 8):
        p = dict(WQ); p['n_bits'] = bits
        q = UniformAffineQuantizer(**p)
        y = q(w)
        out.update({f"w{bits}_y": npy(y), f"w{bits}_delta": npy(q.delta), f"w{bits}_zp": npy(q.zero_point)})
        if bits == 4:
            q.set_inited(True)
            ada = AdaRoundQuantizer(uaq=q, round_mode='learned_hard_sigmoid', weight_tensor=w)
            out["ada_alpha0"] = npy(ada.alpha)
            alpha = ada.alpha.data + 0.4 * torch.randn(w.shape, generator=g)
            ada.alpha.data.copy_(alpha)
            ada.soft_targets = True
            ws = ada(w)
            gy = torch.randn(w.shape, generator=g)
            ws.backward(gy)
            ada.soft_targets = False
            out.update(ada_alpha=npy(alpha), ada_soft=npy(ws), ada_hard=npy(ada(w)), ada_gy=npy(gy), ada_galpha=npy(ada.alpha.grad))
    out["w"] = npy(w)
    # QuantModule with split shortcut (1x1 conv on a concat), W4A8
    conv = torch.nn.Conv2d(48, 24, 1)
    torch.manual_seed(5)
    torch.nn.init.normal_(conv.weight, std=0.1); torch.nn.init.normal_(conv.bias, std=0.1)
    qm = QuantModule(conv, WQ, AQ)
    xin = torch.cat([torch.randn(4, 32, 8, 8, generator=g), 2.0 * torch.randn(4, 16, 8, 8, generator=g)], 1)
    qm.set_quant_state(True, True)
    with torch.no_grad():
        y = qm(xin, split=32)
    out.update(split_w=npy(conv.weight), split_b=npy(conv.bias), split_x=npy(xin), split_y=npy(y),
               split_da=np.stack([npy(qm.act_quantizer.delta), npy(qm.act_quantizer_0.delta)]),
               split_za=np.stack([npy(qm.act_quantizer.zero_point), npy(qm.act_quantizer_0.zero_point)]),
               split_dw0=npy(qm.weight_quantizer.delta), split_zw0=npy(qm.weight_quantizer.zero_point),
               split_dw1=npy(qm.weight_quantizer_0.delta), split_zw1=npy(qm.weight_quantizer_0.zero_point))
    # straight-through gradient of the activation quantizer incl. the step-size gradient
    q = UniformAffineQuantizer(**AQ)
    x = (torch.randn(4, 8, 6, 6, generator=g) * 2).requires_grad_(True)
    q(x); q.set_inited(True)
    q.delta = torch.nn.Parameter(torch.tensor(q.delta) * 0.5)   # force clipping
    gy = torch.randn(x.shape, generator=g)
    q(x).backward(gy)
    out.update(ste_x=npy(x), ste_gy=npy(gy), ste_gx=npy(x.grad), ste_gdelta=npy(q.delta.grad), ste_delta=npy(q.delta), ste_zp=npy(q.zero_point))
    np.savez_compressed(os.path.join(OUT, "unit.npz"), **out)
    print("unit.npz", len(out))


def unit_max_vectors():
    """scale_method 'max' / 'max_scale' (init_quantization_scale_2, reference quant_layer.py:278-345): the constructor default of
    UniformAffineQuantizer, unused by the scripts; (delta, zero_point) and the fake-quantized tensor for every branch."""
    g = torch.Generator().manual_seed(21)
    xs = {"act4d": torch.randn(4, 12, 6, 6, generator=g) * 1.7 + 0.3,
          "pos3d": torch.softmax(torch.randn(3, 16, 16, generator=g), -1),
          "w4d": torch.randn(10, 6, 3, 3, generator=g) * 0.07,
          "w2d": torch.randn(9, 20, generator=g) * 0.2,
          "tiny": torch.randn(5, 7, generator=g) * 1e-10}
    cases = [("act4d", 8, False, False, False, "max"), ("act4d", 8, True, False, False, "max"), ("act4d", 6, False, False, False, "max_scale"),
             ("pos3d", 8, False, False, True, "max"), ("pos3d", 8, False, False, False, "max"),
             ("w4d", 4, True, True, False, "max"), ("w4d", 4, False, True, False, "max"), ("w4d", 8, True, True, False, "max_scale"),
             ("w2d", 4, False, True, False, "max"), ("w2d", 3, True, True, False, "max"), ("tiny", 8, False, False, False, "max")]
    out = {k: npy(v) for k, v in xs.items()}
    out["cases"] = np.array(["|".join(map(str, c)) for c in cases])
    for i, (name, bits, sym, cw, az, method) in enumerate(cases):
        q = UniformAffineQuantizer(n_bits=bits, symmetric=sym, channel_wise=cw, scale_method=method, always_zero=az)
        delta, zp = q.init_quantization_scale_2(xs[name], cw)
        q.delta, q.zero_point = delta, zp
        q.set_inited(True)
        y = q(xs[name])
        zp_t = zp if torch.is_tensor(zp) else torch.tensor(float(zp))
        out.update({f"c{i}_delta": npy(delta), f"c{i}_zp": npy(zp_t.float()), f"c{i}_y": npy(y)})
    # forward() of an un-inited quantizer with the constructor default takes the same route (:251, :260)
    q = UniformAffineQuantizer(n_bits=8)
    out["default_y"] = npy(q(xs["act4d"]))
    # asymmetric, two-sided 'mse' ranges: perform_2D_search (:120-147), 100 clipping widths x n_levels zero-points
    s2d = {"t4": (torch.randn(6, 40, generator=g) * 0.8 + 0.4, 4, False), "t8": (torch.randn(64, generator=g) * 1.5 - 0.2, 8, False),
           "c3": (torch.randn(5, 4, 3, 3, generator=g) * 0.3 + 0.05, 3, True)}
    for key, (x, bits, cw) in s2d.items():
        q = UniformAffineQuantizer(n_bits=bits, symmetric=False, channel_wise=cw, scale_method='mse')
        y = q(x)
        assert q.one_side_dist == 'no'
        out.update({f"s2d_{key}_x": npy(x), f"s2d_{key}_y": npy(y), f"s2d_{key}_delta": npy(q.delta), f"s2d_{key}_zp": npy(q.zero_point)})
    np.savez_compressed(os.path.join(OUT, "unit_max.npz"), **out)
    print("unit_max.npz", len(out))


def _init_all(qnn, cali, bs):
    set_weight_quantize_params(qnn, cali)
    # the generic driver does not reset the LDM matmul quantizers: do it by hand (set_quantize_params_LDM.py:31-36)
    extra = []
    for m in qnn.modules():
        if isinstance(m, QuantQKMatMul):
            extra += [m.act_quantizer_q, m.act_quantizer_k]
        if isinstance(m, QuantSMVMatMul):
            extra += [m.act_quantizer_v, m.act_quantizer_w]
        if isinstance(m, QuantBasicTransformerBlock):
            for a in (m.attn1, m.attn2):
                extra += [a.act_quantizer_q, a.act_quantizer_k, a.act_quantizer_v, a.act_quantizer_w]
    for q in extra:
        q.set_inited(False)
    set_act_quantize_params(qnn, cali, batch_size=bs)
    for q in extra:
        q.set_inited(True)


def _record_losses(module):
    """wrap LossFunction.__call__ of a reference recon module so the loss trace can be read back"""
    trace = []
    orig = module.LossFunction.__call__

    def call(self, pred, tgt, grad=None):
        v = orig(self, pred, tgt, grad)
        trace.append(float(v))
        return v
    module.LossFunction.__call__ = call
    return trace, lambda: setattr(module.LossFunction, "__call__", orig)


def ddim_tiny():
    from ddim.models.diffusion import Model
    cfg = ns(dict(data=dict(image_size=16, channels=3),
                  model=dict(type='simple', in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2], num_res_blocks=1,
                             attn_resolutions=[8], dropout=0.0, resamp_with_conv=True),
                  diffusion=dict(num_diffusion_timesteps=1000)))
    torch.manual_seed(0)
    model = Model(cfg).eval()
    state = {k: npy(v) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(32, 3, 16, 16, generator=g)
    t = torch.randint(0, 1000, (32,), generator=g)
    with torch.no_grad():
        y_fp = model(x[:4], t[:4])
    qnn = QuantModel(model, WQ, AQ, sm_abit=8).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    qnn.model.config.split_shortcut = True
    _init_all(qnn, (x, t), 16)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        y_q = qnn(x[:4], t[:4])
    qnn.set_quant_state(True, False)
    with torch.no_grad():
        y_w = qnn(x[:4], t[:4])
    out = dict(x=npy(x), t=npy(t), y_fp=npy(y_fp), y_w4a8=npy(y_q), y_w4=npy(y_w))
    out.update({"sd." + k: v for k, v in state.items()})
    pack_table("q.", quantizer_table(qnn), out)

    # --- reconstruction traces: deterministic (prob = 1, input_prob = 1), FBR on ---
    kwargs = dict(cali_data=(x, t), iters=4, batch_size=8, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                  act_quant=True, opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=1.0, keep_gpu=True,
                  recon_w=True, recon_a=True, add_loss=0.8)
    random.seed(77); torch.manual_seed(77)
    trace, undo = _record_losses(ref_block_recon)
    blk = qnn.model.down[0].block[0]
    ref_block_recon.block_reconstruction(qnn, blk, **kwargs)
    undo()
    out.update(recon_block_loss=np.array(trace), recon_block_alpha=npy(blk.conv1.weight_quantizer.alpha),
               recon_block_delta=np.array([float(blk.conv1.act_quantizer.delta), float(blk.temb_proj.act_quantizer.delta),
                                           float(blk.conv2.act_quantizer.delta)]))
    random.seed(78); torch.manual_seed(78)
    trace, undo = _record_losses(ref_layer_recon)
    lyr = qnn.model.down[0].downsample.conv
    ref_layer_recon.layer_reconstruction(qnn, lyr, **kwargs)
    undo()
    out.update(recon_layer_loss=np.array(trace), recon_layer_alpha=npy(lyr.weight_quantizer.alpha),
               recon_layer_delta=np.array([float(lyr.act_quantizer.delta)]))
    # attention block recon (QuantAttnBlock: q/k/v/w deltas + 4 convs)
    random.seed(79); torch.manual_seed(79)
    trace, undo = _record_losses(ref_block_recon)
    ab = qnn.model.down[1].attn[0]
    ref_block_recon.block_reconstruction(qnn, ab, **kwargs)
    undo()
    out.update(recon_attn_loss=np.array(trace),
               recon_attn_delta=np.array([float(ab.act_quantizer_q.delta), float(ab.act_quantizer_k.delta),
                                          float(ab.act_quantizer_v.delta), float(ab.act_quantizer_w.delta)]))
    # output after these three units were reconstructed (hard rounding + learned step sizes)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        out["y_after_recon"] = npy(qnn(x[:4], t[:4]))
    np.savez_compressed(os.path.join(OUT, "ddim_tiny.npz"), **out)
    print("ddim_tiny.npz", sum(v.size for v in state.values()), "params")


def _ldm(name, unet_kwargs, ctx_dim=None):
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    torch.manual_seed(0)
    model = UNetModel(**unet_kwargs).eval()
    gi = torch.Generator().manual_seed(3)
    for p in model.parameters():   # un-zero the zero_module convs
        if p.dim() > 1 and float(p.abs().max()) == 0.0:
            p.data.copy_(torch.randn(p.shape, generator=gi) * 0.05)
    state = {k: npy(v) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    res = unet_kwargs["image_size"]
    x = torch.randn(16, unet_kwargs["in_channels"], res, res, generator=g)
    t = torch.randint(0, 1000, (16,), generator=g)
    cali = [x, t]
    if ctx_dim:
        cali.append(torch.randn(16, 3, ctx_dim, generator=g))
    with torch.no_grad():
        y_fp = model(*[c[:4] for c in cali])
    qnn = QuantModel(model, WQ, AQ, sm_abit=8).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    qnn.model.split_shortcut = True
    _init_all(qnn, cali, 8)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        y_q = qnn(*[c[:4] for c in cali])
    out = dict(x=npy(x), t=npy(t), y_fp=npy(y_fp), y_w4a8=npy(y_q))
    if ctx_dim:
        out["ctx"] = npy(cali[2])
    out.update({"sd." + k: v for k, v in state.items()})
    pack_table("q.", quantizer_table(qnn), out)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, sum(v.size for v in state.values()), "params")


def ldm_tiny():
    # church-style: legacy multi-head attention, scale-shift norm, resblock up/down
    _ldm("ldm_tiny.npz", dict(image_size=8, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1, 2],
                              num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_scale_shift_norm=True,
                              resblock_updown=True))
    # bedroom-style: num_head_channels, conv resampling
    _ldm("ldm_tiny_b.npz", dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[2],
                                num_res_blocks=1, channel_mult=[1, 2], num_head_channels=16))


def ldm_xattn_tiny():
    _ldm("ldm_xattn_tiny.npz", dict(image_size=8, in_channels=3, out_channels=3, model_channels=32,
                                    attention_resolutions=[1, 2], num_res_blocks=1, channel_mult=[1, 2], num_heads=2,
                                    use_spatial_transformer=True, transformer_depth=1, context_dim=24), ctx_dim=24)


def cfg_recon():
    """qdiff_control (classifier-free-guidance) reconstruction on the tiny spatial-transformer UNet."""
    import qdiff_control.block_recon as ref_cfg_recon
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    kw = dict(image_size=8, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[1, 2], num_res_blocks=1,
              channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True, transformer_depth=1, context_dim=24)
    torch.manual_seed(0)
    model = UNetModel(**kw).eval()
    gi = torch.Generator().manual_seed(3)
    for p_ in model.parameters():
        if p_.dim() > 1 and float(p_.abs().max()) == 0.0:
            p_.data.copy_(torch.randn(p_.shape, generator=gi) * 0.05)
    state = {k: npy(v) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(999)
    n = 16
    x = torch.randn(n, 3, 8, 8, generator=g)
    t = torch.randint(0, 1000, (n,), generator=g)
    index = torch.zeros(n, dtype=torch.long)
    cond = torch.randn(n, 3, 24, generator=g)
    uncond = torch.randn(1, 3, 24, generator=g).repeat(n, 1, 1)
    cali = (x, t, index, cond, uncond)
    qnn = QuantModel(model, WQ, AQ, sm_abit=8).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    qnn.model.split_shortcut = True
    _init_all(qnn, [torch.cat([x, x]), torch.cat([t, t]), torch.cat([uncond, cond])], 8)
    out = dict(x=npy(x), t=npy(t), index=npy(index), cond=npy(cond), uncond=npy(uncond))
    out.update({"sd." + k: v for k, v in state.items()})
    pack_table("q.", quantizer_table(qnn), out)
    kwargs = dict(cali_data=cali, iters=4, batch_size=4, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2, act_quant=True,
                  opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=1.0, keep_gpu=True, recon_w=True, recon_a=True,
                  add_loss=0.8)
    random.seed(55); torch.manual_seed(55)
    trace, undo = _record_losses(ref_cfg_recon)
    res = qnn.model.input_blocks[1][0]
    ref_cfg_recon.block_reconstruction(qnn, res, **kwargs)
    undo()
    out.update(recon_res_loss=np.array(trace), recon_res_alpha=npy(res.in_layers[2].weight_quantizer.alpha))
    random.seed(56); torch.manual_seed(56)
    trace, undo = _record_losses(ref_cfg_recon)
    tb = qnn.model.input_blocks[1][1].transformer_blocks[0]
    ref_cfg_recon.block_reconstruction(qnn, tb, **kwargs)
    undo()
    out.update(recon_tb_loss=np.array(trace),
               recon_tb_delta=np.array([float(tb.attn1.act_quantizer_q.delta), float(tb.attn1.act_quantizer_w.delta),
                                        float(tb.attn2.act_quantizer_k.delta), float(tb.attn2.act_quantizer_v.delta)]))
    np.savez_compressed(os.path.join(OUT, "cfg_xattn_tiny.npz"), **out)
    print("cfg_xattn_tiny.npz")


def api_goldens():
    """Reference drivers that round 1 left untested: AttnBlock_layer_reconstruction (qdiff/attn_layer_recon.py:13-133) and the
    whole-model walk recon_block_Qmodel.recon() (qdiff/recon_block_Qmodel.py:18-94) on the tiny DDIM UNet of ddim_tiny()."""
    import qdiff.attn_layer_recon as ref_attn_recon
    from qdiff import recon_block_Qmodel
    from ddim.models.diffusion import Model
    cfg = ns(dict(data=dict(image_size=16, channels=3),
                  model=dict(type='simple', in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2], num_res_blocks=1,
                             attn_resolutions=[8], dropout=0.0, resamp_with_conv=True),
                  diffusion=dict(num_diffusion_timesteps=1000)))
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(32, 3, 16, 16, generator=g)
    t = torch.randint(0, 1000, (32,), generator=g)

    def fresh(c=None):
        torch.manual_seed(0)
        qnn = QuantModel(Model(c or cfg).eval(), WQ, AQ, sm_abit=8).eval()
        qnn.set_first_last_layer_to_8bit()
        qnn.disable_network_output_quantization()
        qnn.model.config.split_shortcut = True
        _init_all(qnn, (x, t), 16)
        return qnn
    kwargs = dict(cali_data=(x, t), iters=4, batch_size=8, weight=0.01, asym=True, b_range=(20, 2), warmup=0.2,
                  act_quant=True, opt_mode='mse', lr_a=4e-4, lr_w=1e-2, p=2.0, input_prob=1.0, keep_gpu=True,
                  recon_w=True, recon_a=True, add_loss=0.8)
    out = {}
    qnn = fresh()
    random.seed(81); torch.manual_seed(81)
    trace, undo = _record_losses(ref_attn_recon)
    ab = qnn.model.down[1].attn[0]
    ref_attn_recon.AttnBlock_layer_reconstruction(qnn, ab, **kwargs)
    undo()
    out.update(attn_layer_loss=np.array(trace),
               attn_layer_delta=np.array([float(ab.act_quantizer_q.delta), float(ab.act_quantizer_k.delta),
                                          float(ab.act_quantizer_v.delta), float(ab.act_quantizer_w.delta)]))
    # whole-model walk, 2 iterations per unit (the reference hard-codes two block/attention pairs on the attention level,
    # recon_block_Qmodel.py:33-38, so this UNet has num_res_blocks = 2)
    cfg2 = ns(dict(data=dict(image_size=16, channels=3),
                   model=dict(type='simple', in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2], num_res_blocks=2,
                              attn_resolutions=[8], dropout=0.0, resamp_with_conv=True),
                   diffusion=dict(num_diffusion_timesteps=1000)))
    qnn = fresh(cfg2)
    out.update({"walk_sd." + k: npy(v) for k, v in qnn.model.state_dict().items() if "quantizer" not in k and "org_" not in k})
    pack_table("walk_q.", quantizer_table(qnn), out)
    kw2 = dict(kwargs); kw2.update(iters=2)
    random.seed(82); torch.manual_seed(82)
    tb, undo_b = _record_losses(ref_block_recon)
    tl, undo_l = _record_losses(ref_layer_recon)
    order = []
    ob, ol = ref_block_recon.block_reconstruction, ref_layer_recon.layer_reconstruction
    names = {id(m): n for n, m in qnn.named_modules()}
    ref_driver = sys.modules["qdiff.recon_block_Qmodel"]      # (the package attribute of that name is the class)

    def wrap(fn, kind, trace):
        def call(model, unit, **k):
            n0 = len(trace)
            r = fn(model, unit, **k)
            order.append((kind, names[id(unit)], [float(v) for v in trace[n0:]]))
            return r
        return call
    ref_driver.block_reconstruction = wrap(ob, "block", tb)
    ref_driver.layer_reconstruction = wrap(ol, "layer", tl)
    recon_block_Qmodel(None, qnn, (x, t), kw2).recon()
    ref_driver.block_reconstruction, ref_driver.layer_reconstruction = ob, ol
    undo_b(); undo_l()
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        y = qnn(x[:4], t[:4])
    out.update(walk_kinds=np.array([k for k, _, _ in order]), walk_names=np.array([n for _, n, _ in order]),
               walk_losses=np.array([l for _, _, l in order]), walk_y=npy(y))
    np.savez_compressed(os.path.join(OUT, "ddim_tiny_api.npz"), **out)
    print("ddim_tiny_api.npz", len(order), "units:", [n for _, n, _ in order])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    which = sys.argv[1:] or ["unit", "unit_max", "ddim", "ldm", "xattn", "cfg", "api"]
    if "unit" in which:
        unit_vectors()
    if "unit_max" in which:
        unit_max_vectors()
    if "ddim" in which:
        ddim_tiny()
    if "ldm" in which:
        ldm_tiny()
    if "xattn" in which:
        ldm_xattn_tiny()
    if "cfg" in which:
        cfg_recon()
    if "api" in which:
        api_goldens()
