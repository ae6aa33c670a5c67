"""-m gpu: the reference-facing drivers round 1 left untested -- AttnBlock_layer_reconstruction, recon_block_Qmodel.recon()
(goldens recorded from the unmodified reference, oracle/make_golden.py api_goldens) and the sampler-driven scale-init drivers
set_*_quantize_params_LDM / _Conditional / _Stable (the reference's samplers are out of scope, SURVEY.md section 2 row 20: a stub
reproduces their single-step `quant_unet=True` mode, ldm/models/diffusion/ddim.py:101-106, :186-216)."""
import random
import sys
import types

import numpy as np
import pytest
import torch

import helpers as H
from test_gpu_model import _product, _set_split_ddim, _set_split_ldm, RECON_KW

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def test_attn_block_layer_reconstruction_trace(cuda):
    """qdiff/attn_layer_recon.py:13-133: only the q/k/v/softmax step sizes of a QuantAttnBlock are trained"""
    from qdiff.attn_layer_recon import AttnBlock_layer_reconstruction
    g, api = H.load("ddim_tiny.npz"), H.load("ddim_tiny_api.npz")
    qnn = _product(g, H.ddim_tiny_model(), cuda, _set_split_ddim)
    x, t = T(g["x"]).to(cuda), T(g["t"]).to(cuda)
    with torch.no_grad():
        qnn(x[:4], t[:4])
    H.install_qparams(qnn, H.qtable(g))
    ab = qnn.model.down[1].attn[0]
    alpha_free = [m.weight_quantizer for m in (ab.q, ab.k, ab.v, ab.proj_out)]
    random.seed(81); torch.manual_seed(81)
    losses = AttnBlock_layer_reconstruction(qnn, ab, cali_data=(x, t), return_losses=True, **RECON_KW)
    ref = api["attn_layer_loss"]
    assert abs(losses[0].item() - ref[0]) <= 2e-3 * abs(ref[0])
    assert np.allclose(losses.cpu().numpy(), ref, rtol=5e-2)
    d = [float(q.delta) for q in (ab.act_quantizer_q, ab.act_quantizer_k, ab.act_quantizer_v, ab.act_quantizer_w)]
    assert np.allclose(d, api["attn_layer_delta"], rtol=2e-2, atol=2e-4)
    assert all(not hasattr(wq, "alpha") for wq in alpha_free)          # weights are not touched (attn_layer_recon.py:42-58)


def test_recon_block_Qmodel_walk(cuda):
    """recon_block_Qmodel.recon() visits the units in the reference's order (qdiff/recon_block_Qmodel.py:26-89: unrolled attention
    levels, `up` walked backwards) with the reference's unit kinds, and reproduces its first loss values"""
    from unet_zoo.ddpm_unet import DDPMUNet
    from qdiff import QuantModel, recon_block_Qmodel
    api = H.load("ddim_tiny_api.npz")
    model = DDPMUNet(ch=32, out_ch=3, ch_mult=(1, 2, 2), num_res_blocks=2, attn_resolutions=(8,), dropout=0.0, in_channels=3,
                     resolution=16, resamp_with_conv=True).eval()
    model.load_state_dict({k[len("walk_sd."):]: T(api[k]) for k in api.files if k.startswith("walk_sd.")})
    qnn = QuantModel(model.to(cuda), H.WQ, H.AQ, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    _set_split_ddim(qnn.model)
    g = H.load("ddim_tiny.npz")
    x, t = T(g["x"]).to(cuda), T(g["t"]).to(cuda)
    with torch.no_grad():
        qnn(x[:4], t[:4])
    names = [str(n) for n in api["walk_q.names"]]
    table = {n: (T(api[f"walk_q.{i}_delta"]), T(api[f"walk_q.{i}_zp"]), int(api[f"walk_q.{i}_bits"])) for i, n in enumerate(names)}
    H.install_qparams(qnn, table)
    driver = sys.modules["qdiff.recon_block_Qmodel"]
    order = []
    mod_names = {id(m): n for n, m in qnn.named_modules()}
    ob, ol = driver.block_reconstruction, driver.layer_reconstruction

    def wrap(fn, kind):
        def call(model_, unit, **k):
            losses = fn(model_, unit, return_losses=True, **k)
            order.append((kind, mod_names[id(unit)], losses.cpu().numpy()))
            return losses
        return call
    driver.block_reconstruction, driver.layer_reconstruction = wrap(ob, "block"), wrap(ol, "layer")
    kw = dict(RECON_KW); kw.update(iters=2, cali_data=(x, t))
    random.seed(82); torch.manual_seed(82)
    try:
        out = recon_block_Qmodel(None, qnn, (x, t), kw).recon()
    finally:
        driver.block_reconstruction, driver.layer_reconstruction = ob, ol
    assert out is qnn
    assert [(k, n) for k, n, _ in order] == list(zip([str(k) for k in api["walk_kinds"]], [str(n) for n in api["walk_names"]]))
    ref = api["walk_losses"]
    for i in range(3):                                   # units in front of the first cross-platform code flip: tight
        assert np.allclose(order[i][2], ref[i], rtol=2e-2), (order[i][1], order[i][2], ref[i])
    got = np.array([l[0] for _, _, l in order])
    assert np.all(np.isfinite(got))
    # later units see inputs that went through earlier (chaotic) units: same order of magnitude, unit by unit
    ratio = got / ref[:, 0]
    assert np.median(np.abs(np.log(ratio))) < 0.25, ratio
    with torch.no_grad():
        y = qnn(x[:4], t[:4])
    assert torch.isfinite(y).all() and H.rel_l2(y.cpu(), T(api["walk_y"])) < 0.5


# ---- sampler-driven scale init ------------------------------------------------------------------------------------------------
class _FakeLatentDiffusion(torch.nn.Module):
    """the two attributes the drivers touch: `.model.diffusion_model` (the QuantModel) and the conditioning hooks"""

    def __init__(self, qnn, ctx_dim=None, tokens=1):
        super().__init__()
        self.model = torch.nn.Module()
        self.model.diffusion_model = qnn
        self.cond_stage_key = "class_label"
        self.ctx_dim, self.tokens = ctx_dim, tokens

    @property
    def device(self):
        return next(self.parameters()).device

    def get_learned_conditioning(self, c):
        n = len(c) if isinstance(c, list) else len(next(iter(c.values())))
        seed = 7 if (isinstance(c, list) and c and c[0] == "") or (isinstance(c, dict) and int(next(iter(c.values()))[0]) == 1000) else 11
        g = torch.Generator().manual_seed(seed)
        return torch.randn(1, self.tokens, self.ctx_dim, generator=g).repeat(n, 1, 1).to(self.device)

    def apply_model(self, x, t, c):
        return self.model.diffusion_model(x, t, c)


class _StubSampler:
    """`sample(..., quant_unet=True, cali_data=...)`: one UNet call on the calibration slice, with the classifier-free-guidance
    batch [x;x], [t;t], [uc;c] when a guidance scale is set (reference ddim.py:101-106, 191-210)"""
    calls = []

    def __init__(self, model):
        self.model = model

    def sample(self, S=None, batch_size=None, shape=None, conditioning=None, unconditional_guidance_scale=1., unconditional_conditioning=None,
               quant_unet=False, cali_data=None, **kw):
        assert quant_unet and cali_data is not None
        x, t = cali_data[0], cali_data[1]
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            args = (x, t, conditioning)
        else:
            args = (torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([unconditional_conditioning, conditioning]))
        _StubSampler.calls.append(tuple(a.detach().clone() if torch.is_tensor(a) else a for a in args))
        return self.model.apply_model(*args), None


@pytest.fixture
def stub_samplers(monkeypatch):
    for name in ("ldm", "ldm.models", "ldm.models.diffusion"):
        if name not in sys.modules:
            monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    for sub, cls in (("ddim", "DDIMSampler"), ("ddim_control", "DDIMSampler_control"), ("plms", "PLMSSampler")):
        m = types.ModuleType(f"ldm.models.diffusion.{sub}")
        setattr(m, cls, _StubSampler)
        monkeypatch.setitem(sys.modules, f"ldm.models.diffusion.{sub}", m)
    _StubSampler.calls = []
    return _StubSampler


def _tables_equal(a, b):
    ta = {n: m for n, m in a.named_modules() if hasattr(m, "delta") and getattr(m, "delta", None) is not None and hasattr(m, "n_levels")}
    tb = {n: m for n, m in b.named_modules() if hasattr(m, "delta") and getattr(m, "delta", None) is not None and hasattr(m, "n_levels")}
    assert set(ta) == set(tb) and len(ta) > 20
    for n in ta:
        assert torch.equal(ta[n].delta.detach(), tb[n].delta.detach()), n
        assert torch.equal(ta[n].zero_point, tb[n].zero_point), n
        assert ta[n].inited is True and tb[n].inited is True, n


def test_set_quantize_params_LDM_driver(cuda, stub_samplers):
    from qdiff import set_weight_quantize_params, set_act_quantize_params
    from qdiff.set_quantize_params_LDM import set_weight_quantize_params_LDM, set_act_quantize_params_LDM
    g = H.load("ldm_tiny.npz")
    x, t = T(g["x"])[:8].to(cuda), T(g["t"])[:8].to(cuda)
    index = torch.zeros(8, dtype=torch.long, device=cuda)
    args = types.SimpleNamespace(custom_steps=20, eta=0.0)
    q1 = _product(g, H.ldm_model("ldm_tiny.npz"), cuda, _set_split_ldm)
    ld = _FakeLatentDiffusion(q1).to(cuda)
    set_weight_quantize_params_LDM(ld, (x, t, index), args)
    set_act_quantize_params_LDM(ld, (x, t, index), args, batch_size=4)
    assert len(stub_samplers.calls) == 3                      # one weight pass (8 samples), two activation batches of 4
    q2 = _product(g, H.ldm_model("ldm_tiny.npz"), cuda, _set_split_ldm)
    set_weight_quantize_params(q2, (x, t))
    set_act_quantize_params(q2, (x, t), batch_size=4, all_attention=True)
    _tables_equal(q1, q2)


@pytest.mark.parametrize("which", ["Conditional", "Stable"])
def test_set_quantize_params_cfg_drivers(cuda, stub_samplers, which):
    from qdiff import set_weight_quantize_params, set_act_quantize_params
    import qdiff_control
    g = H.load("ldm_xattn_tiny.npz")
    n = 4
    x, t = T(g["x"])[:n].to(cuda), T(g["t"])[:n].to(cuda)
    index = torch.zeros(n, dtype=torch.long, device=cuda)
    q1 = _product(g, H.ldm_model("ldm_xattn_tiny.npz"), cuda, _set_split_ldm)
    ld = _FakeLatentDiffusion(q1, ctx_dim=24, tokens=3).to(cuda)
    if which == "Conditional":
        args = types.SimpleNamespace(custom_steps=20, ddim_eta=0.0, scale=3.0, data=torch.arange(n))
        w_fn, a_fn = qdiff_control.set_weight_quantize_params_Conditional, qdiff_control.set_act_quantize_params_Conditional
    else:
        args = types.SimpleNamespace(custom_steps=20, ddim_eta=0.0, scale=7.5, list_prompts=["a church"] * n, plms=False, C=3, H=64, W=64, f=8)
        w_fn, a_fn = qdiff_control.set_weight_quantize_params_Stable, qdiff_control.set_act_quantize_params_Stable
    w_fn(ld, (x, t, index), args)
    a_fn(ld, (x, t, index), args, batch_size=2)
    calls = list(stub_samplers.calls)
    assert len(calls) == 3 and calls[0][0].shape[0] == 4 and calls[1][0].shape[0] == 4      # CFG doubles every slice (2 -> 4)
    # the same calibration batches pushed through the plain drivers give the same tables
    q2 = _product(g, H.ldm_model("ldm_xattn_tiny.npz"), cuda, _set_split_ldm)
    set_weight_quantize_params(q2, calls[0])
    act = [torch.cat([calls[1][i], calls[2][i]]) for i in range(3)]
    set_act_quantize_params(q2, act, batch_size=4, all_attention=True)
    _tables_equal(q1, q2)


def test_packed_w4_export_roundtrip(cuda, tmp_path):
    """f4: export after reconstruction (hard AdaRound codes, learned step sizes) -> file -> fresh model: bit-identical forward;
    4-bit codes cost half a byte per weight"""
    from qdiff.block_recon import block_reconstruction
    from qdiff.export import export_packed, load_packed, packed_nbytes
    g = H.load("ddim_tiny.npz")
    qnn = _product(g, H.ddim_tiny_model(), cuda, _set_split_ddim)
    x, t = T(g["x"]).to(cuda), T(g["t"]).to(cuda)
    with torch.no_grad():
        qnn(x[:4], t[:4])
    H.install_qparams(qnn, H.qtable(g))
    random.seed(1); torch.manual_seed(1)
    block_reconstruction(qnn, qnn.model.down[0].block[0], cali_data=(x, t), **dict(RECON_KW, iters=3))
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        y0 = qnn(x[:8], t[:8])
    path = tmp_path / "w4a8.pt"
    blob = export_packed(qnn, str(path))
    n_w = sum(m.weight.numel() for m in qnn.modules() if hasattr(m, "weight_quantizer"))
    n_w4 = sum(m.weight.numel() for m in qnn.modules() if hasattr(m, "weight_quantizer") and m.weight_quantizer.n_bits <= 4)
    code_bytes = sum(p["codes"].numel() for e in blob["layers"].values() for p in e["weights"])
    assert code_bytes <= n_w4 // 2 + (n_w - n_w4) + 64 * len(blob["layers"])          # two 4-bit codes per byte
    assert packed_nbytes(blob) < 0.3 * 4 * sum(p.numel() for p in H.ddim_tiny_model().parameters())
    fresh = _product(g, H.ddim_tiny_model(), cuda, _set_split_ddim)
    with torch.no_grad():
        torch.manual_seed(123)
        for p in fresh.parameters():
            p.add_(torch.randn_like(p) * 0.01)                 # whatever the fresh model holds is overwritten
        fresh(x[:4], t[:4])                                    # creates the split twins
    load_packed(fresh, str(path))
    fresh.set_quant_state(True, True)
    with torch.no_grad():
        y1 = fresh(x[:8], t[:8])
    assert torch.equal(y0, y1)
    assert sum(v == "int8" for v in fresh.path_report().values()) >= len(fresh.path_report()) - 1
