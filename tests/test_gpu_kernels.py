"""-m gpu parity tests of the CUDA kernels (through the C ABI) against the CPU oracle.

Integer codes: bit-exact.  Floating-point outputs: tolerance stated at each assert.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import qdiff_oracle as O

pytestmark = pytest.mark.gpu


def _act_params(x, n_bits=8):
    d, z, _ = O.init_scale(x.cpu(), n_bits, channel_wise=False, sym=True)
    return d.reshape(1), z.reshape(1)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(4, 64, 16, 16), (3, 5, 7), (1,), (2, 320, 33)])
@pytest.mark.parametrize("zp", [128.0, 127.0, 0.0])
def test_uaq_fwd_codes_bit_exact(cuda, shape, zp):
    from edadm import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g) * 2.0
    delta = torch.tensor([0.0173])
    z = torch.tensor([zp])
    y, codes = ops.uaq_forward(x.to(cuda), delta.to(cuda), z.to(cuda), 256, want_codes=True)
    ref_codes = O.uaq_codes(x, delta, z, 256)
    ref_y = O.uaq_forward(x, delta, z, 256)
    assert torch.equal(codes.cpu().float(), ref_codes)
    assert torch.equal(y.cpu(), ref_y)  # same fp32 ops in the same order -> bit-exact


def test_uaq_fwd_channelwise_weights(cuda):
    from edadm import ops
    g = torch.Generator().manual_seed(1)
    w = torch.randn(48, 32, 3, 3, generator=g) * 0.1
    d, z, _ = O.init_scale(w, 4, channel_wise=True)
    y, codes = ops.uaq_forward(w.to(cuda), d.to(cuda), z.to(cuda), 16, want_codes=True)
    assert torch.equal(codes.cpu().float(), O.uaq_codes(w, d, z, 16))
    assert torch.equal(y.cpu(), O.uaq_forward(w, d, z, 16))


@pytest.mark.parametrize("prob", [1.0, 0.5])
def test_uaq_bwd_matches_autograd(cuda, prob):
    from edadm import ops
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(8, 32, 16, 16, generator=g) * 3.0)
    delta = torch.tensor(0.02)
    z = torch.tensor(128.0)
    keep = (torch.rand(x.shape, generator=g) < prob) if prob < 1 else None
    gy = torch.randn(x.shape, generator=g)
    # oracle: autograd through the restated forward
    xr = x.clone().requires_grad_(True)
    dr = delta.clone().requires_grad_(True)
    O.uaq_forward(xr, dr, z, 256, keep).backward(gy)
    xc = x.to(cuda).requires_grad_(True)
    dc = delta.to(cuda).requires_grad_(True)
    y = ops.uaq_fake_quant(xc, dc, z.to(cuda), 256, keep.to(cuda) if keep is not None else None)
    y.backward(gy.to(cuda))
    assert torch.equal(xc.grad.cpu(), xr.grad)  # pass-through / zero: exact
    # step-size gradient: fp64-accumulated sum of ~65k fp32 terms vs torch's fp32 reduction: rel 1e-4
    assert abs(dc.grad.item() - dr.grad.item()) <= 1e-4 * max(1.0, abs(dr.grad.item()))


def test_uaq_philox_mask_consistent_and_rate(cuda):
    from edadm import ops
    x = torch.randn(1 << 20, device=cuda)
    d = torch.tensor([0.05], device=cuda)
    z = torch.tensor([128.0], device=cuda)
    xq = ops.uaq_forward(x, d, z, 256)
    y1 = ops.uaq_forward(x, d, z, 256, prob=0.5, seed=7, offset=0)
    y2 = ops.uaq_forward(x, d, z, 256, prob=0.5, seed=7, offset=0)
    assert torch.equal(y1, y2)
    kept = (y1 == xq) & (xq != x)
    dropped = (y1 == x) & (xq != x)
    frac = kept.sum().item() / max(1, (kept | dropped).sum().item())
    assert abs(frac - 0.5) < 0.01
    # backward draws the same mask
    xg = x.clone().requires_grad_(True)
    ops.uaq_fake_quant(xg, d, z, 256, None, 0.5, 7, 0).sum().backward()
    inr = (torch.round(x / d) + z).clamp(0, 255) == (torch.round(x / d) + z)
    expect = torch.where(y1 == x, torch.ones_like(x), inr.float())
    same = (xg.grad == expect) | (xq == x)
    assert bool(same.all())


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("soft", [True, False])
def test_adaround_fwd_bwd(cuda, soft):
    from edadm import ops
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 48, 3, 3, generator=g) * 0.05
    d, z, _ = O.init_scale(w, 4, channel_wise=True)
    alpha = O.adaround_init_alpha(w, d) + 0.3 * torch.randn(w.shape, generator=g)
    a_gpu = ops.adaround_init_alpha(w.to(cuda), d.to(cuda))
    # alpha init: logf/div ulp differences between libm and CUDA: abs 1e-4 on values of O(1..10)
    assert torch.allclose(a_gpu.cpu(), O.adaround_init_alpha(w, d), atol=1e-4, rtol=1e-5)
    ar = alpha.clone().requires_grad_(True)
    ref = O.adaround_forward(w, ar, d, z, 16, soft)
    ac = alpha.to(cuda).requires_grad_(True)
    out = ops.adaround_fake_quant(w.to(cuda), ac, d.to(cuda), z.to(cuda), 16, soft)
    if soft:
        # sigmoid via expf: 1-2 ulp vs torch CPU; values are O(delta*8): atol 1e-6
        assert torch.allclose(out.cpu(), ref, atol=1e-6, rtol=1e-5)
        gy = torch.randn(w.shape, generator=g)
        ref.backward(gy)
        out.backward(gy.to(cuda))
        assert torch.allclose(ac.grad.cpu(), ar.grad, atol=1e-7, rtol=1e-4)
    else:
        assert torch.equal(out.detach().cpu(), ref.detach())
        _, codes = ops.adaround_forward(w.to(cuda), alpha.to(cuda), d.to(cuda), z.to(cuda), 16, False, want_codes=True)
        assert torch.equal(codes.cpu().float(), O.adaround_codes(w, alpha, d, z, 16))


def test_round_reg(cuda):
    from edadm import ops
    g = torch.Generator().manual_seed(4)
    alpha = torch.randn(20000, generator=g) * 2
    ar = alpha.clone().requires_grad_(True)
    ref = O.round_reg(ar, 8.0, 0.01)
    ref.backward()
    ac = alpha.to(cuda).requires_grad_(True)
    out = ops.round_reg(ac, 8.0, 0.01)
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert torch.allclose(ac.grad.cpu(), ar.grad, atol=1e-6, rtol=1e-3)


@pytest.mark.parametrize("p", [2.0, 2.4])
@pytest.mark.parametrize("shape", [(32, 64, 16, 16), (4, 77, 320), (2, 3)])
def test_lp_loss(cuda, p, shape):
    from edadm import ops
    g = torch.Generator().manual_seed(5)
    a = torch.randn(shape, generator=g)
    b = torch.randn(shape, generator=g)
    ar = a.clone().requires_grad_(True)
    ref = O.lp_loss(ar, b, p)
    ref.backward()
    ac = a.to(cuda).requires_grad_(True)
    out = ops.lp_loss(ac, b.to(cuda), p)
    out.backward()
    # fp64 accumulation on the GPU vs fp32 pairwise sum in torch: rel 1e-5
    assert abs(out.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.allclose(ac.grad.cpu(), ar.grad, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------------------------------
def _ref_int_conv(codes_a, za, wcodes, zw, stride=1, padding=0):
    """exact integer convolution of (qa - za) with (qw - zw[n]) in float64"""
    a = codes_a.double() - za
    w = wcodes.double() - zw.reshape(-1, 1, 1, 1).double()
    return F.conv2d(a, w, None, stride=stride, padding=padding)


@pytest.mark.parametrize("B,C,H,W,pad", [(2, 64, 16, 16, 1), (3, 20, 8, 8, 1), (1, 3, 32, 32, 1), (2, 130, 4, 4, 0),
                                         # large enough for the TMA-staged producer: full tiles, ragged channel tile,
                                         # ragged pixel tile (HW % 32 != 0), narrow heads (CT = 32 / 64)
                                         (8, 256, 32, 32, 1), (6, 200, 32, 32, 1), (40, 128, 6, 10, 1), (64, 24, 32, 32, 0),
                                         (24, 64, 24, 24, 1)])
def test_act_quant_nhwc_codes(cuda, B, C, H, W, pad):
    from edadm import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, C, H, W, generator=g) * 1.5
    d, z = _act_params(x)
    aq = ops.ActQuant(d.to(cuda), z.to(cuda), 256)
    q, chsum = ops.act_quant_nhwc(x.to(cuda), aq, pad, want_chsum=True)
    ref = O.uaq_codes(x, d, z, 256).permute(0, 2, 3, 1)
    q = q.cpu()
    assert torch.equal(q[:, pad:pad + H, pad:pad + W, :C].float(), ref)
    assert int(q[..., C:].sum()) == 0
    if pad:
        assert bool((q[:, 0, :, :C] == int(z.item())).all()) and bool((q[:, :, -1, :C] == int(z.item())).all())
    assert torch.equal(chsum.cpu().long(), q.long().sum(-1))


def test_act_quant_nhwc_tma_split_and_strided(cuda):
    """TMA-staged producer: split quantizers (quant_layer.py:415-419), prescale, and a batch-strided source view"""
    from edadm import ops
    g = torch.Generator().manual_seed(61)
    B, C, H, W, split = 8, 192, 32, 32, 128
    x = torch.randn(B, C, H, W, generator=g) * 1.2
    x[:, split:] *= 3.0
    d0, z0 = _act_params(x[:, :split])
    d1, z1 = _act_params(x[:, split:])
    aq = ops.ActQuant(d0.to(cuda), z0.to(cuda), 256, split, d1.to(cuda), z1.to(cuda), 256)
    q, chsum = ops.act_quant_nhwc(x.to(cuda), aq, 1, want_chsum=True)
    ref = torch.cat([O.uaq_codes(x[:, :split], d0, z0, 256), O.uaq_codes(x[:, split:], d1, z1, 256)], 1).permute(0, 2, 3, 1)
    q = q.cpu()
    assert torch.equal(q[:, 1:-1, 1:-1, :C].float(), ref)
    assert torch.equal(chsum.cpu().long(), q.long().sum(-1))
    # q / k / v style view: 3 tensors interleaved along channels, read in place with prescale
    big = torch.randn(B, 3 * C, H, W, generator=g)
    view = big.to(cuda)[:, C:2 * C]
    d, z = _act_params(big[:, C:2 * C] * 0.25)
    qv, _ = ops.act_quant_nhwc(view, ops.ActQuant(d.to(cuda), z.to(cuda), 256, prescale=0.25), 0)
    refv = O.uaq_codes(big[:, C:2 * C] * 0.25, d, z, 256).permute(0, 2, 3, 1)
    assert torch.equal(qv.cpu()[..., :C].float(), refv)


@pytest.mark.parametrize("M,K", [(200, 128), (64, 77), (4096, 320), (1, 512)])
def test_act_quant_rows_codes(cuda, M, K):
    from edadm import ops
    g = torch.Generator().manual_seed(7)
    x = torch.randn(M, K, generator=g)
    d, z = _act_params(x)
    q, rs = ops.act_quant_rows(x.to(cuda), ops.ActQuant(d.to(cuda), z.to(cuda), 256), want_rowsum=True)
    ref = O.uaq_codes(x, d, z, 256)
    assert torch.equal(q.cpu()[:, :K].float(), ref)
    assert torch.equal(rs.cpu().long(), ref.long().sum(1))


@pytest.mark.parametrize("n_bits", [4, 8])
def test_pack_weight_codes(cuda, n_bits):
    from edadm import ops
    g = torch.Generator().manual_seed(8)
    w = torch.randn(40, 24, 3, 3, generator=g) * 0.07
    d, z, _ = O.init_scale(w, n_bits, channel_wise=True)
    L = 2 ** n_bits
    pw = ops.pack_weight(w.to(cuda), d.to(cuda), z.to(cuda), L, want_codes=True, w4=False)
    ref = O.uaq_codes(w, d, z, L)
    assert torch.equal(pw.codes.cpu().float(), ref)
    if n_bits == 4:
        # nibble-packed storage: [Np][taps][Cp/2], byte j of every 4-byte word = code[c0+j] | code[c0+4+j] << 4
        p4 = ops.pack_weight(w.to(cuda), d.to(cuda), z.to(cuda), L, want_codes=True, w4=True)
        assert p4.w4 and p4.Cp == 32 and p4.wq.dtype == torch.uint8 and tuple(p4.wq.shape) == (p4.Np, 9, 16)
        assert torch.equal(p4.codes.cpu().float(), ref)
        b = p4.wq.cpu()[:40].reshape(40, 9, 4, 4).long()                      # [n][tap][word][byte]
        lo, hi = b & 15, b >> 4
        un = torch.stack([lo, hi], dim=3).reshape(40, 9, 32)                    # word -> codes c0..c0+3, c0+4..c0+7
        codes_nc = ref.reshape(40, 24, 9).permute(0, 2, 1).long()               # [n][tap][c]
        assert torch.equal(un[..., :24], codes_nc)
        assert torch.equal(un[..., 24:], z.reshape(40, 1, 1).long().expand(40, 9, 8))   # padding == zero-point code
        assert torch.equal(p4.zoff.cpu()[:40].long(), z.reshape(-1).long())
        assert torch.equal(p4.wsum_eff.cpu()[:40].long(), (codes_nc - z.reshape(40, 1, 1).long()).sum((1, 2)))
    zoff = z.reshape(-1) if L <= 128 else torch.full((40,), 128.0)
    wq = pw.wq.cpu()[:40].reshape(40, 3, 3, -1)[..., :24].permute(0, 3, 1, 2).float()
    assert torch.equal(wq, ref - zoff.reshape(-1, 1, 1, 1))
    # alpha path (hard AdaRound codes)
    alpha = O.adaround_init_alpha(w, d) + 0.5 * torch.randn(w.shape, generator=g)
    pw2 = ops.pack_weight(w.to(cuda), d.to(cuda), z.to(cuda), L, alpha=alpha.to(cuda), want_codes=True)
    assert torch.equal(pw2.codes.cpu().float(), O.adaround_codes(w, alpha, d, z, L))


# ------------------------------------------------------------------------------------------------
def _run_qconv(cuda, x, w, bias, n_bits_w, pad, kind="conv"):
    """full integer path: quantize activations, pack weights, tcgen05 GEMM"""
    from edadm import ops
    d_a, z_a = _act_params(x)
    d_w, z_w, _ = O.init_scale(w, n_bits_w, channel_wise=True)
    Lw = 2 ** n_bits_w
    pw = ops.pack_weight(w.to(cuda), d_w.to(cuda), z_w.to(cuda), Lw)
    aq = ops.ActQuant(d_a.to(cuda), z_a.to(cuda), 256)
    if kind == "conv":
        B, C, H, W = x.shape
        q, chsum = ops.act_quant_nhwc(x.to(cuda), aq, pad, want_chsum=pw.needs_rowsum)
        R, S = w.shape[2], w.shape[3]
        Ho, Wo = H + 2 * pad - R + 1, W + 2 * pad - S + 1
        rs = ops.conv_rowsum(chsum, Ho, Wo, R, S, 1) if pw.needs_rowsum else None
        out = torch.empty(B, w.shape[0], Ho, Wo, device=cuda)
        ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, out, Ho * Wo, bias=None if bias is None else bias.to(cuda), rowsum=rs)
    else:
        q, rs = ops.act_quant_rows(x.to(cuda), aq, want_rowsum=pw.needs_rowsum)
        out = torch.empty(x.shape[0], w.shape[0], device=cuda)
        ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, out, 1, bias=None if bias is None else bias.to(cuda), rowsum=rs)
    torch.cuda.synchronize()
    # oracle: reference fake-quant forward (quant_layer.py:406-437) in fp32 on CPU
    act_q = [dict(delta=d_a, zero_point=z_a, n_levels=256)]
    w_q = [dict(delta=d_w, zero_point=z_w, n_levels=Lw)]
    if kind == "conv":
        ref = O.quant_module_forward(x, w, bias, "conv2d", dict(stride=1, padding=pad), act_q, w_q)
        exact = _ref_int_conv(O.uaq_codes(x, d_a, z_a, 256), z_a.item(), O.uaq_codes(w, d_w, z_w, Lw), z_w.reshape(-1), 1, pad)
        exact = exact * (d_a.double() * d_w.reshape(1, -1, 1, 1).double())
        if bias is not None:
            exact = exact + bias.reshape(1, -1, 1, 1).double()
    else:
        ref = O.quant_module_forward(x, w, bias, "linear", {}, act_q, w_q)
        exact = (O.uaq_codes(x, d_a, z_a, 256).double() - z_a.item()) @ (O.uaq_codes(w, d_w, z_w, Lw).double() - z_w.reshape(-1, 1).double()).t()
        exact = exact * (d_a.double() * d_w.reshape(1, -1).double())
        if bias is not None:
            exact = exact + bias.reshape(1, -1).double()
    return out.cpu(), ref, exact.float()


def _rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("M,K,N,bits", [(256, 128, 64, 4), (300, 320, 320, 4), (128, 512, 1280, 4), (77, 768, 96, 8),
                                        (1000, 96, 24, 8), (4096, 384, 3072, 4)])
def test_qgemm_linear(cuda, M, K, N, bits):
    g = torch.Generator().manual_seed(9)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    out, ref, exact = _run_qconv(cuda, x, w, bias, bits, 0, kind="linear")
    # integer accumulation is exact; the only rounding is the final fp32 scale+bias: 1e-6 of the exact value
    assert _rel_l2(out, exact) < 1e-6
    # vs the reference fp32 fake-quant forward (north_star tolerance 1e-3; fp32 summation order only)
    assert _rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("B,C,H,N,k", [(3, 64, 16, 128, 3), (5, 96, 8, 100, 3), (2, 32, 32, 48, 1)])
def test_qgemm_bias_img_epilogue(cuda, B, C, H, N, k):
    """`conv(x) + emb[:, :, None, None]` folded into the epilogue == the separate broadcast add (quant_block.py:112-113)"""
    from edadm import ops
    g = torch.Generator().manual_seed(36)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(N, C, k, k, generator=g) * 0.05
    bias = torch.randn(N, generator=g).to(cuda)
    emb = torch.randn(B, N, 1, 1, generator=g).to(cuda)
    d_a, z_a = _act_params(x)
    d_w, z_w, _ = O.init_scale(w, 4, channel_wise=True)
    pw = ops.pack_weight(w.to(cuda), d_w.to(cuda), z_w.to(cuda), 16)
    aq = ops.ActQuant(d_a.to(cuda), z_a.to(cuda), 256)
    q, _ = ops.act_quant_nhwc(x.to(cuda), aq, k // 2, cp=pw.Cp)
    plain = ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, torch.empty(B, N, H, H, device=cuda), H * H, bias=bias)
    fused = ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, torch.empty(B, N, H, H, device=cuda), H * H, bias=bias, bias_img=emb)
    assert torch.equal(fused, plain + emb)


@pytest.mark.parametrize("shape,N,k", [((2, 128, 16, 16), 128, 3), ((4, 48, 8, 8), 100, 3), ((2, 200, 16, 16), 96, 1),
                                       ((300, 320), 320, 0), ((77, 720), 98, 0), ((16, 80, 32, 32), 256, 3)])
def test_qgemm_w4_storage_matches_s8(cuda, shape, N, k):
    """weights kept as 4-bit codes (two per byte) and unpacked in shared memory == the same GEMM on s8 codes, bit for bit"""
    from edadm import ops
    g = torch.Generator().manual_seed(35)
    x = torch.randn(*shape, generator=g)
    conv = len(shape) == 4
    w = torch.randn(N, shape[1], k, k, generator=g) * 0.05 if conv else torch.randn(N, shape[1], generator=g) * 0.05
    bias = torch.randn(N, generator=g).to(cuda)
    d_a, z_a = _act_params(x)
    d_w, z_w, _ = O.init_scale(w, 4, channel_wise=True)
    aq = ops.ActQuant(d_a.to(cuda), z_a.to(cuda), 256)
    outs = []
    for w4 in (False, True):
        pw = ops.pack_weight(w.to(cuda), d_w.to(cuda), z_w.to(cuda), 16, w4=w4)
        assert pw.w4 == w4
        if conv:
            q, _ = ops.act_quant_nhwc(x.to(cuda), aq, k // 2, cp=pw.Cp)
            oshape, out_hw = (shape[0], N, shape[2], shape[3]), shape[2] * shape[3]
        else:
            q, _ = ops.act_quant_rows(x.to(cuda), aq)
            oshape, out_hw = (shape[0], N), 1
        outs.append(ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, torch.empty(oshape, device=cuda), out_hw, bias=bias))
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("shape,N,k,bits", [((2, 128, 16, 16), 128, 3, 4), ((4, 64, 8, 8), 100, 3, 4), ((2, 64, 16, 16), 96, 3, 8),
                                            ((300, 320), 320, 0, 4), ((77, 768), 98, 0, 4), ((200, 128), 64, 0, 8)])
def test_qgemm_residual_epilogue(cuda, shape, N, k, bits):
    """`conv(x) + residual` fused into the epilogue == the separate fp32 add, bit for bit (quant_block.py:116,192)"""
    from edadm import ops
    g = torch.Generator().manual_seed(33)
    x = torch.randn(*shape, generator=g)
    conv = len(shape) == 4
    w = torch.randn(N, shape[1], k, k, generator=g) * 0.05 if conv else torch.randn(N, shape[1], generator=g) * 0.05
    bias = torch.randn(N, generator=g).to(cuda)
    d_a, z_a = _act_params(x)
    d_w, z_w, _ = O.init_scale(w, bits, channel_wise=True)
    pw = ops.pack_weight(w.to(cuda), d_w.to(cuda), z_w.to(cuda), 2 ** bits)
    aq = ops.ActQuant(d_a.to(cuda), z_a.to(cuda), 256)
    if conv:
        q, chsum = ops.act_quant_nhwc(x.to(cuda), aq, k // 2, want_chsum=pw.needs_rowsum)
        rs = ops.conv_rowsum(chsum, shape[2], shape[3], k, k, 1) if pw.needs_rowsum else None
        oshape, out_hw = (shape[0], N, shape[2], shape[3]), shape[2] * shape[3]
    else:
        q, rs = ops.act_quant_rows(x.to(cuda), aq, want_rowsum=pw.needs_rowsum)
        oshape, out_hw = (shape[0], N), 1
    res = torch.randn(*oshape, generator=g).to(cuda)
    plain = ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, torch.empty(oshape, device=cuda), out_hw, bias=bias, rowsum=rs)
    fused = ops.qgemm_i8(q, pw, aq.delta0, aq.zp0, torch.empty(oshape, device=cuda), out_hw, bias=bias, rowsum=rs, residual=res)
    assert torch.equal(fused, plain + res)


@pytest.mark.parametrize("B,C,H,N,k,bits", [(2, 128, 16, 128, 3, 4), (1, 224, 64, 224, 3, 4), (4, 64, 8, 96, 3, 4),
                                            (8, 256, 4, 256, 3, 4), (2, 3, 32, 128, 3, 8), (2, 128, 32, 3, 3, 8),
                                            (2, 192, 16, 384, 1, 4), (32, 32, 2, 48, 3, 4)])
def test_qgemm_conv(cuda, B, C, H, N, k, bits):
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(N, C, k, k, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    out, ref, exact = _run_qconv(cuda, x, w, bias, bits, k // 2, kind="conv")
    assert _rel_l2(out, exact) < 1e-6
    assert _rel_l2(out, ref) < 1e-5


# ------------------------------------------------------------------------------------------------
def _attn_quant_params(q, k, v, probs):
    out = {}
    for name, t, kw in (("q", q, dict(sym=True)), ("k", k, dict(sym=True)), ("v", v, dict(sym=True)), ("w", probs, dict(sym=False))):
        d, z, _ = O.init_scale(t, 8, False, running={}, **kw)
        out[name] = dict(delta=d, zero_point=z, n_levels=256)
    return out


def _to_aquant(cuda, qp):
    from edadm import ops
    return ops.AttnQuant(*[(qp[n]["delta"].to(cuda), qp[n]["zero_point"].to(cuda), 256) for n in ("q", "k", "v", "w")])


@pytest.mark.parametrize("B,C,H", [(2, 256, 16), (3, 64, 8), (2, 32, 4), (1, 96, 12)])
def test_qattn_ddim_layout(cuda, B, C, H):
    """QuantAttnBlock attention core (quant_block.py:431-445): fused kernel vs oracle.  q/k/v codes are bit-exact by
    construction (same producer kernels as above); the probabilities go through a different-order fp32 softmax so single
    codes may flip at .5 boundaries -> relative L2 <= 1e-3 on the block output (north_star tolerance)."""
    from edadm import ops
    g = torch.Generator().manual_seed(21)
    q, k, v = (torch.randn(B, C, H, H, generator=g) * s for s in (1.0, 1.2, 0.8))
    scale = int(C) ** -0.5
    fp_probs = torch.softmax(torch.bmm(q.reshape(B, C, -1).permute(0, 2, 1), k.reshape(B, C, -1)) * scale, dim=2)
    qp = _attn_quant_params(q, k, v, fp_probs)
    ref = O.attn_core_ddim(q, k, v, qp)
    out = ops.qattn_bct(q.reshape(B, C, -1).to(cuda), k.reshape(B, C, -1).to(cuda), v.reshape(B, C, -1).to(cuda),
                        _to_aquant(cuda, qp), 1.0, scale)
    assert _rel_l2(out.cpu().reshape(ref.shape), ref) < 1e-3


@pytest.mark.parametrize("B,heads,ch,T", [(2, 8, 24, 1024), (2, 2, 16, 64), (1, 4, 48, 256), (3, 7, 32, 100), (2, 8, 96, 16)])
def test_qattn_ldm_legacy_layout(cuda, B, heads, ch, T):
    from edadm import ops
    g = torch.Generator().manual_seed(22)
    qkv = torch.randn(B, heads * 3 * ch, T, generator=g)
    q, k, v = qkv.reshape(B * heads, ch * 3, T).split(ch, dim=1)
    scale = 1 / (ch ** 0.25)
    fp_probs = torch.softmax(torch.einsum("bct,bcs->bts", q * scale, k * scale), -1)
    qp = _attn_quant_params(q * scale, k * scale, v, fp_probs)
    ref = O.attn_core_ldm(qkv, heads, qp)
    out = ops.qattn_bct(q.to(cuda), k.to(cuda), v.to(cuda), _to_aquant(cuda, qp), scale, 1.0)
    assert _rel_l2(out.cpu().reshape(ref.shape), ref) < 1e-3


@pytest.mark.parametrize("B,heads,d,Tq,Tk", [(2, 1, 384, 256, 256), (2, 8, 40, 128, 77), (1, 1, 576, 64, 1), (2, 8, 80, 300, 300),
                                             (1, 1, 960, 64, 64), (1, 1, 384, 1024, 1024), (2, 1, 576, 256, 256), (1, 1, 300, 200, 200),
                                             (1, 1, 512, 128, 128)])
def test_qattn_cross_layout(cuda, B, heads, d, Tq, Tk):
    from edadm import ops
    g = torch.Generator().manual_seed(23)
    q = torch.randn(B, Tq, heads * d, generator=g)
    k = torch.randn(B, Tk, heads * d, generator=g)
    v = torch.randn(B, Tk, heads * d, generator=g)
    scale = d ** -0.5

    def sp(t):
        return t.reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(t.shape[0] * heads, t.shape[1], d)
    fp_probs = (torch.einsum("bid,bjd->bij", sp(q), sp(k)) * scale).softmax(-1)
    qp = _attn_quant_params(sp(q), sp(k), sp(v), fp_probs)
    ref = O.attn_core_cross(q, k, v, heads, scale, qp)
    out = ops.qattn_bnd(sp(q).contiguous().to(cuda), sp(k).contiguous().to(cuda), sp(v).contiguous().to(cuda), heads,
                        _to_aquant(cuda, qp), scale)
    assert _rel_l2(out.cpu(), ref) < 1e-3


def test_qattn_sd_self_attention_max_keys(cuda):
    """Stable-Diffusion self-attention at the 64x64 level (quant_block.py:204-235: T = 4096 tokens, d = 40) -- exactly the
    kernel's key limit (csrc/qattn.cu ATT_MAX_KEYS), two of the eight heads.  Ranges are set by hand (max-abs for q/k/v, a
    clipping range for the probabilities as the mse search picks) so that the CPU oracle finishes in seconds."""
    from edadm import ops
    heads, d, T = 2, 40, 4096
    g = torch.Generator().manual_seed(29)
    q, k, v = (torch.randn(1, T, heads * d, generator=g) for _ in range(3))
    scale = d ** -0.5

    def sp(t):
        return t.reshape(1, T, heads, d).permute(0, 2, 1, 3).reshape(heads, T, d)
    qp = {n: dict(delta=t.abs().max() / 127.5, zero_point=torch.tensor(128.), n_levels=256) for n, t in (("q", q), ("k", k), ("v", v))}
    p_max = float((torch.einsum("bid,bjd->bij", sp(q), sp(k)) * scale).softmax(-1).max())
    qp["w"] = dict(delta=torch.tensor(0.25 * p_max / 255), zero_point=torch.tensor(0.), n_levels=256)
    ref = O.attn_core_cross(q, k, v, heads, scale, qp)
    out = ops.qattn_bnd(sp(q).contiguous().to(cuda), sp(k).contiguous().to(cuda), sp(v).contiguous().to(cuda), heads,
                        _to_aquant(cuda, qp), scale)
    assert _rel_l2(out.cpu(), ref) < 1e-3


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,H,N", [(2, 192, 32, 3), (3, 64, 13, 4), (1, 30, 40, 1), (2, 128, 8, 2)])
def test_conv3x3_small_n(cuda, B, C, H, N):
    """output-layer stencil (fp32 x fake-quantized W8) vs the exact fp64 convolution; fp32 summation order only"""
    from edadm import ops
    g = torch.Generator().manual_seed(44)
    x = torch.randn(B, C, H, H + 3, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g) * 0.05
    bias = torch.randn(N, generator=g)
    out = ops.conv3x3_small_n(x.to(cuda), w.to(cuda), bias.to(cuda)).cpu()
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    assert _rel_l2(out, ref) < 1e-6
    # GroupNorm + SiLU of the output head folded into the patch load
    if C % 2 == 0:
        gn = torch.nn.GroupNorm(2, C).to(cuda)
        with torch.no_grad():
            gn.weight.copy_(torch.randn(C, generator=g).to(cuda) * 0.3 + 1)
            gn.bias.copy_(torch.randn(C, generator=g).to(cuda) * 0.2)
            a, s = ops.gn_fold(x.to(cuda), gn.weight, gn.bias, 2, gn.eps)
            fused = ops.conv3x3_small_n(x.to(cuda), w.to(cuda), bias.to(cuda), affine=(a, s, True))
            ref2 = F.conv2d(F.silu(gn(x.to(cuda))).double(), w.to(cuda).double(), bias.to(cuda).double(), padding=1).float()
        assert _rel_l2(fused, ref2) < 1e-5
    assert ops.conv3x3_small_n_ok(x.to(cuda), w.to(cuda), dict(stride=(1, 1), padding=(1, 1), dilation=(1, 1), groups=1))
    assert not ops.conv3x3_small_n_ok(x.to(cuda), w.to(cuda), dict(stride=(2, 2), padding=(1, 1), dilation=(1, 1), groups=1))


@pytest.mark.parametrize("B,C,H,scale_shift", [(3, 64, 16, False), (2, 192, 8, True), (2, 96, 4, False), (2, 768, 4, True),
                                               (2, 192, 32, False), (1, 64, 64, False)])
def test_groupnorm_silu_quant_producer(cuda, B, C, H, scale_shift):
    """GroupNorm (+ scale-shift) + SiLU + quantize in one pass vs the module-by-module path: the codes agree except where
    silu(gn(x))/delta lands within an ulp of a .5 boundary (different-order fp32 arithmetic in GroupNorm itself)."""
    from edadm import ops
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, C, H, H, generator=g) * 2 + 0.3
    gn = torch.nn.GroupNorm(32, C, eps=1e-5)
    gn.weight.data = torch.randn(C, generator=g) * 0.5 + 1
    gn.bias.data = torch.randn(C, generator=g) * 0.2
    scale = torch.randn(B, C, 1, 1, generator=g) * 0.3 if scale_shift else None
    shift = torch.randn(B, C, 1, 1, generator=g) * 0.3 if scale_shift else None
    with torch.no_grad():
        h = gn(x)
        if scale_shift:
            h = h * (1 + scale) + shift
        h = F.silu(h)
    d, z = _act_params(h)
    ref_codes = O.uaq_codes(h, d, z, 256).permute(0, 2, 3, 1)
    gn = gn.to(cuda)
    a, s = ops.gn_fold(x.to(cuda), gn.weight, gn.bias, 32, gn.eps, None if scale is None else scale.to(cuda),
                       None if shift is None else shift.to(cuda))
    q, _ = ops.norm_act_quant_nhwc(x.to(cuda), a, s, True, ops.ActQuant(d.to(cuda), z.to(cuda), 256), 1)
    codes = q.cpu()[:, 1:-1, 1:-1, :C].float()
    diff = (codes - ref_codes).abs()
    assert float(diff.max()) <= 1.0                                   # never more than one code step
    assert float((diff > 0).float().mean()) < 2e-3                    # and only at rounding boundaries (CPU GroupNorm / SiLU differ from CUDA's)
    # against the SAME modules run by torch on this GPU the producer uses ATen-CUDA's SiLU arithmetic (x / (1 + expf(-x)), IEEE
    # division), so only last-ulp differences of the GroupNorm statistics (fp64 sums here, fp32 Welford there) can move a code
    with torch.no_grad():
        hg = gn(x.to(cuda))
        if scale_shift:
            hg = hg * (1 + scale.to(cuda)) + shift.to(cuda)
        hg = F.silu(hg)
        gpu_codes = torch.clamp(torch.round(hg / d.to(cuda)) + z.to(cuda), 0, 255).permute(0, 2, 3, 1).cpu()
    flips = float(((codes - gpu_codes).abs() > 0).float().mean())
    print(f"flip rate vs torch-CUDA modules: {flips:.2e}")
    assert flips < (2e-4 if scale_shift else 3e-5)                    # scale-shift: h*(1+scale)+shift is folded into the affine (one FMA)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K", [(300, 384), (64, 96), (1000, 1536), (33, 50), (2048, 576), (512, 960)])
def test_layernorm_quant_rows(cuda, M, K):
    """LayerNorm + activation quantizer in one pass vs nn.LayerNorm followed by the quantizer (rounding-boundary flips only)"""
    from edadm import ops
    g = torch.Generator().manual_seed(51)
    x = torch.randn(M, K, generator=g) * 2 + 0.3
    ln = torch.nn.LayerNorm(K)
    ln.weight.data = torch.randn(K, generator=g) * 0.3 + 1
    ln.bias.data = torch.randn(K, generator=g) * 0.2
    with torch.no_grad():
        y = ln(x)
    d, z = _act_params(y)
    ref = O.uaq_codes(y, d, z, 256)
    q, rs = ops.layernorm_quant_rows(x.to(cuda), ln.weight.to(cuda), ln.bias.to(cuda), ln.eps,
                                     ops.ActQuant(d.to(cuda), z.to(cuda), 256), want_rowsum=True)
    codes = q.cpu()[:, :K].float()
    diff = (codes - ref).abs()
    assert float(diff.max()) <= 1.0 and float((diff > 0).float().mean()) < 2e-3
    assert int(q.cpu()[:, K:].sum()) == 0
    assert torch.equal(rs.cpu().long(), q.cpu().long().sum(1))
    # against torch's own LayerNorm on this GPU (same final FMA; only last-ulp differences of mean / rstd can move a code)
    with torch.no_grad():
        yg = ln.to(cuda)(x.to(cuda))
        gpu_codes = torch.clamp(torch.round(yg / d.to(cuda)) + z.to(cuda), 0, 255).cpu()
    flips = float(((codes - gpu_codes).abs() > 0).float().mean())
    print(f"LayerNorm producer flip rate vs torch-CUDA: {flips:.2e}")
    assert flips < 5e-5


@pytest.mark.parametrize("M,K", [(200, 1536), (77, 96), (1024, 3072), (5, 20)])
def test_geglu_quant_rows(cuda, M, K):
    """GEGLU gate + quantizer in one pass == F.gelu on the device followed by the quantizer (same erf formulation)"""
    from edadm import ops
    g = torch.Generator().manual_seed(52)
    h = torch.randn(M, 2 * K, generator=g) * 1.5
    a, gate = h.to(cuda).chunk(2, dim=-1)
    y = (a * F.gelu(gate)).cpu()
    d, z = _act_params(y)
    ref = O.uaq_codes(y, d, z, 256)
    q, rs = ops.geglu_quant_rows(h.to(cuda), ops.ActQuant(d.to(cuda), z.to(cuda), 256), want_rowsum=True)
    codes = q.cpu()[:, :K].float()
    diff = (codes - ref).abs()
    assert float(diff.max()) <= 1.0 and float((diff > 0).float().mean()) < 1e-4
    assert torch.equal(rs.cpu().long(), q.cpu().long().sum(1))


def test_transformer_block_fusions_match_unfused(cuda):
    """QuantBasicTransformerBlock + SpatialTransformer with LayerNorm / GEGLU / residual / token-layout fusions vs the same
    modules run one by one (backend.fuse_norm off): same integer GEMMs, only normalisation rounding differs"""
    import copy
    from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
    from qdiff.quant_layer import backend, QuantModule
    from unet_zoo.ldm_unet import UNetModel
    torch.manual_seed(7)
    model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
                      attention_resolutions=(1, 2), channel_mult=(1, 2), num_heads=4, use_spatial_transformer=True,
                      transformer_depth=1, context_dim=48).to(cuda).eval()
    for p in model.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)            # zero-initialised proj_out would hide the path
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 4, 16, 16, generator=g).to(cuda)
    t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
    ctx = torch.randn(8, 5, 48, generator=g).to(cuda)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        set_weight_quantize_params(qnn, (x, t, ctx))
        set_act_quantize_params(qnn, (x, t, ctx), all_attention=True)
        backend.fuse_norm = True
        y1 = qnn(x, t, ctx)
        paths = {n: m.last_path for n, m in qnn.named_modules() if isinstance(m, QuantModule)}
        backend.fuse_norm = False
        try:
            y0 = qnn(x, t, ctx)
        finally:
            backend.fuse_norm = True
    assert sum(p == 'int8' for p in paths.values()) >= len(paths) - 2
    assert _rel_l2(y1, y0) < 5e-2     # chaotic amplification of single code flips; typically ~1e-3
    assert torch.isfinite(y1).all()


def _inited_module(cuda, org, sample):
    """QuantModule around `org` with weight / activation scales searched on `sample` and the integer path armed"""
    from qdiff.quant_layer import QuantModule
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qm = QuantModule(org, wq, aq).to(cuda).eval()
    qm.set_quant_state(True, True)
    with torch.no_grad():
        qm(sample)                         # un-inited quantizers run their search on the first forward
    for q in (qm.weight_quantizer, qm.act_quantizer):
        q.set_inited(True)
    return qm


def test_token_layout_and_geglu_paths_are_exact(cuda):
    """tokens_out / forward_from_tokens / forward_geglu change layouts and producers, not arithmetic: bit-identical outputs"""
    from qdiff.quant_layer import backend
    torch.manual_seed(11)
    B, C, H, W, N = 4, 64, 16, 16, 96
    x = torch.randn(B, C, H, W, device=cuda)
    gn = torch.nn.GroupNorm(32, C, eps=1e-6).to(cuda)
    with torch.no_grad():
        xn = gn(x)
        proj_in = _inited_module(cuda, torch.nn.Conv2d(C, N, 1), xn)
        y_nchw = proj_in.forward_prenorm(x, gn, silu=False)
        y_tok = proj_in.forward_prenorm(x, gn, silu=False, tokens_out=True)
        assert proj_in.last_path == 'int8' and tuple(y_tok.shape) == (B, H * W, N) and y_tok.is_contiguous()
        assert torch.equal(y_tok, y_nchw.flatten(2).permute(0, 2, 1))
        tok = torch.randn(B, H * W, N, device=cuda)
        proj_out = _inited_module(cuda, torch.nn.Conv2d(N, C, 1), tok.permute(0, 2, 1).reshape(B, N, H, W))
        ref = proj_out(tok.permute(0, 2, 1).reshape(B, N, H, W).contiguous()) + x
        got = proj_out.forward_from_tokens(tok, (H, W), residual=x)
        assert proj_out.last_path == 'int8' and torch.equal(got, ref)
        h = torch.randn(B, 50, 2 * 128, device=cuda) * 1.5
        a, g = h.chunk(2, dim=-1)
        lin = _inited_module(cuda, torch.nn.Linear(128, 64), a * F.gelu(g))
        res = torch.randn(B, 50, 64, device=cuda)
        fused = lin.forward_geglu(h, residual=res)
        backend.fuse_norm = False
        try:
            plain = lin.forward_geglu(h, residual=res)
        finally:
            backend.fuse_norm = True
        assert torch.equal(fused, plain)


def test_resample_producers(cuda):
    """resampling ResBlock chain: pooled GroupNorm+SiLU pass and nearest upsampling on codes vs the torch modules"""
    from edadm import ops
    g = torch.Generator().manual_seed(71)
    B, C, H = 3, 64, 16
    x = (torch.randn(B, C, H, H, generator=g) * 1.3).to(cuda)
    gn = torch.nn.GroupNorm(32, C).to(cuda)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(C, generator=g).to(cuda) * 0.3 + 1)
        gn.bias.copy_(torch.randn(C, generator=g).to(cuda) * 0.2)
        y = F.silu(gn(x))
        a, s = ops.gn_fold(x, gn.weight, gn.bias, 32, gn.eps)
        pooled = ops.norm_act_pool2(x, a, s, True)
        assert _rel_l2(pooled.cpu(), F.avg_pool2d(y, 2).cpu()) < 1e-5
        d, z = _act_params(y.cpu())
        aq = ops.ActQuant(d.to(cuda), z.to(cuda), 256)
        q_lo, _ = ops.act_quant_nhwc(y, aq, 0)
        q_hi = ops.upsample2x_codes(q_lo, C, 1, aq)
        ref, _ = ops.act_quant_nhwc(F.interpolate(y, scale_factor=2, mode="nearest"), aq, 1)
        assert torch.equal(q_hi, ref)


# ---- round 2: second-generation GEMM (CTA pairs, TMA-store epilogue, code-emitting epilogues) -------------------------------
@pytest.mark.parametrize("ctas", ["1", "2"])
@pytest.mark.parametrize("B,C,H,N,k,res", [(6, 192, 32, 192, 3, True), (3, 128, 64, 96, 3, False), (40, 64, 8, 320, 3, True),
                                           (70, 96, 4, 160, 3, True), (5, 72, 16, 40, 1, False)])
def test_qgemm2_conv_pairs_and_tma_store(cuda, monkeypatch, ctas, B, C, H, N, k, res):
    """cta_group::2 pairs / single CTAs, NCHW TMA-store epilogue incl. tiles that span several images (HW < 128), M and N tails,
    TMA-loaded residual: exact vs an fp64 convolution of the integer codes"""
    from edadm import ops
    monkeypatch.setenv("EDADM_GEMM_CTAS", ctas)
    monkeypatch.setenv("EDADM_GEMM_V2", "1")
    g = torch.Generator().manual_seed(B * 131 + C)
    x = torch.randn(B, C, H, H, generator=g).to(cuda)
    w = (torch.randn(N, C, k, k, generator=g) * 0.05).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    d = torch.tensor([0.03], device=cuda); z = torch.tensor([128.], device=cuda)
    dw = (w.flatten(1).abs().amax(1) / 7.5).reshape(-1, 1, 1, 1); zw = torch.full_like(dw, 8.)
    pw = ops.pack_weight(w, dw, zw, 16, want_codes=True, w4=False)
    q, _ = ops.act_quant_nhwc(x, ops.ActQuant(d, z, 256), k // 2)
    out = torch.empty(B, N, H, H, device=cuda)
    r = torch.randn(B, N, H, H, generator=g).to(cuda) if res else None
    ops.qgemm_i8(q, pw, d, z, out, H * H, bias=bias, residual=r)
    p = k // 2
    ai = q[:, p:q.shape[1] - p or None, p:q.shape[2] - p or None, :C].permute(0, 3, 1, 2).double() - 128.0
    ref = F.conv2d(ai, pw.codes.double() - 8.0, padding=p) * (0.03 * dw.double().reshape(1, -1, 1, 1)) + bias.double().reshape(1, -1, 1, 1)
    if res:
        ref = ref + r.double()
    assert _rel_l2(out.double(), ref) < 1e-6


@pytest.mark.parametrize("M,K,N,geglu,w_bits", [(300, 384, 384, False, 4), (4096, 384, 3072, True, 4), (1000, 96, 192, True, 4),
                                                (77, 320, 80, False, 8), (513, 128, 64, True, 8),
                                                # large M: the weight-resident schedule (N block fixed per CTA, activation-only ring)
                                                (32768, 384, 384, False, 4), (16384, 384, 3072, True, 4), (32900, 512, 256, False, 4)])
def test_qgemm_codes_epilogue_bit_exact(cuda, M, K, N, geglu, w_bits):
    """the code-emitting epilogue (plain / GEGLU-gated) == fp32 GEMM output followed by the standalone quantizer producers"""
    from edadm import ops
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    L = 2 ** w_bits
    d = torch.tensor([0.03], device=cuda); z = torch.tensor([128.], device=cuda)
    dw = (w.abs().amax(1) / (L / 2 - 0.5)).reshape(-1, 1); zw = torch.full_like(dw, float(L // 2 - (w_bits == 8)))
    pw = ops.pack_weight(w, dw, zw, L, w4=False)
    aq = ops.ActQuant(d, z, 256)
    q, rowsum = ops.act_quant_rows(x, aq, want_rowsum=pw.needs_rowsum)
    y = torch.empty(M, N, device=cuda)
    ops.qgemm_i8(q, pw, d, z, y, 1, bias=bias, rowsum=rowsum)
    cd = torch.tensor([0.011], device=cuda); cz = torch.tensor([128.], device=cuda)
    cons = ops.ActQuant(cd, cz, 256)
    if geglu:
        want, want_rs = ops.geglu_quant_rows(y, cons, want_rowsum=True)
    else:
        want, want_rs = ops.act_quant_rows(y, cons, want_rowsum=True)
    got, got_rs = ops.qgemm_i8_codes(q, pw, d, z, (cd, cz, 256), bias=bias, rowsum=rowsum, geglu=geglu, want_rowsum=True)
    n_out = N // 2 if geglu else N
    assert torch.equal(got[:, :n_out], want[:, :n_out])
    assert torch.equal(got_rs, want_rs)


def test_layernorm_multi_equals_single(cuda):
    from edadm import ops
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(333, 384, generator=g) * 2).to(cuda)
    ln = torch.nn.LayerNorm(384).to(cuda)
    with torch.no_grad():
        ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.3, 0.3)
    aqs = [ops.ActQuant(torch.tensor([s], device=cuda), torch.tensor([zp], device=cuda), 256) for s, zp in ((0.02, 128.), (0.031, 127.), (0.05, 128.))]
    multi = ops.layernorm_quant_rows_multi(x, ln.weight, ln.bias, ln.eps, aqs, [True, False, True])
    for aq, (qc, rs), want_rs in zip(aqs, multi, [True, False, True]):
        q1, r1 = ops.layernorm_quant_rows(x, ln.weight, ln.bias, ln.eps, aq, want_rowsum=True)
        assert torch.equal(qc, q1)
        assert (rs is None) == (not want_rs) and (rs is None or torch.equal(rs, r1))


@pytest.mark.parametrize("ctx_tokens,heads", [(1, 1), (5, 1), (1, 4)])
def test_transformer_epilogue_fusions_are_exact(cuda, ctx_tokens, heads):
    """GEGLU / q / k codes from the GEMM epilogue, the shared LayerNorm pass and the one-key cross-attention shortcut leave a
    QuantBasicTransformerBlock's output bit-identical to the module-by-module integer path (backend.fuse_epilogue off)"""
    from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
    from qdiff.quant_layer import backend
    from unet_zoo.ldm_unet import UNetModel
    torch.manual_seed(17)
    model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
                      attention_resolutions=(1, 2), channel_mult=(1, 2), num_heads=heads, use_spatial_transformer=True,
                      transformer_depth=1, context_dim=48).to(cuda).eval()
    for p in model.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(8, 4, 16, 16, generator=g).to(cuda)
    t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
    ctx = torch.randn(8, ctx_tokens, 48, generator=g).to(cuda)
    qnn.set_quant_state(True, True)
    with torch.no_grad():
        set_weight_quantize_params(qnn, (x, t, ctx))
        set_act_quantize_params(qnn, (x, t, ctx), all_attention=True)
        qnn.set_quant_state(True, True)
        y1 = qnn(x, t, ctx)
        backend.fuse_epilogue = False
        try:
            y0 = qnn(x, t, ctx)
        finally:
            backend.fuse_epilogue = True
    assert torch.isfinite(y1).all()
    assert torch.equal(y1, y0)


@pytest.mark.parametrize("split", [True, False])
def test_lazy_skip_concatenation_is_exact(cuda, split):
    """up path of an LDM UNet with the skip concatenation never materialised (two-source GroupNorm statistics, one producer launch
    per source into channel slices of the code tensor, split quantizers at the boundary) vs th.cat: identical output"""
    from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
    from qdiff.quant_layer import backend
    from edadm import ops
    from unet_zoo.ldm_unet import UNetModel
    torch.manual_seed(29)
    model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=2,
                      attention_resolutions=(), channel_mult=(1, 2, 3), num_heads=1).to(cuda).eval()
    for p in model.parameters():
        if p.dim() > 1 and float(p.detach().abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
    model.split_shortcut = split
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 4, 16, 16, generator=g).to(cuda)
    t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
    made = []
    orig = ops.CatPair.__init__
    with torch.no_grad():
        set_weight_quantize_params(qnn, (x, t))
        set_act_quantize_params(qnn, (x, t), all_attention=True)
        qnn.set_quant_state(True, True)
        ops.CatPair.__init__ = lambda self, a, b: (made.append(1), orig(self, a, b))[1]
        try:
            y1 = qnn(x, t)
        finally:
            ops.CatPair.__init__ = orig
        n_lazy = len(made)
        backend.lazy_cat = False
        try:
            y0 = qnn(x, t)
        finally:
            backend.lazy_cat = True
    assert n_lazy >= 4, n_lazy                 # the up path really took the two-source route
    assert torch.equal(y1, y0)


def test_scale_search_is_reproducible(cuda):
    """the one-pass candidate scoring sums in a fixed order: scores (and the chosen step size) are bit-identical run to run, so a
    quantizer initialised twice on the same tensor gets the same (delta, zero_point)"""
    from edadm import ops
    from qdiff.quant_layer import UniformAffineQuantizer
    g = torch.Generator().manual_seed(41)
    x = (torch.randn(6, 192, 40, 40, generator=g) * 0.8).to(cuda)
    res = []
    for i in range(5):
        junk = torch.empty(1 << (20 + i), device=cuda).normal_()          # different allocator / scheduling state per round
        q = UniformAffineQuantizer(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True)
        q(x)
        res.append((q.delta.detach().clone(), q.zero_point.clone()))
        del junk
    assert all(torch.equal(r[0], res[0][0]) and torch.equal(r[1], res[0][1]) for r in res[1:])
    K = 100
    delta = (torch.linspace(0.2, 1.0, K) * (x.abs().max().item() * 2 / 255)).to(cuda)
    zp = torch.full((K,), 128.0, device=cuda)
    s0 = ops.mse_search_scores(x.reshape(1, -1), delta.reshape(1, K), zp.reshape(1, K), 256, 2.4)
    for _ in range(4):
        assert torch.equal(ops.mse_search_scores(x.reshape(1, -1), delta.reshape(1, K), zp.reshape(1, K), 256, 2.4), s0)


@pytest.mark.parametrize("B,C0,C1,H,N,k", [(8, 128, 64, 16, 96, 1), (4, 384, 192, 32, 192, 1), (130, 64, 64, 8, 160, 1), (6, 96, 32, 16, 64, 3)])
def test_split_shortcut_single_launch_is_exact(cuda, B, C0, C1, H, N, k):
    """edadm_qgemm_i8_split (two K ranges, two TMEM accumulators, one launch) equals the two-launch form (second launch accumulating
    onto the first one's fp32 output) bit for bit, and is the sum of the two ranges' own convolutions"""
    from edadm import ops
    g = torch.Generator().manual_seed(B * 7 + C0)
    x = (torch.randn(B, C0 + C1, H, H, generator=g) * 1.1).to(cuda)
    w = (torch.randn(N, C0 + C1, k, k, generator=g) * 0.05).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    pad = k // 2
    d0, z0 = _act_params(x[:, :C0].cpu())
    d1, z1 = _act_params(x[:, C0:].cpu() * 0.5)
    d0, z0, d1, z1 = (t.to(cuda) for t in (d0, z0, d1, z1))
    aq = ops.ActQuant(d0, z0, 256, C0, d1, z1, 256)
    q, _ = ops.act_quant_nhwc(x, aq, pad)
    packs = []
    for c0, c1 in ((0, C0), (C0, C0 + C1)):
        ws = w[:, c0:c1]
        dw = (ws.flatten(1).abs().amax(1) / 7.5).clamp_min(1e-8).reshape(-1, 1, 1, 1)
        packs.append(ops.pack_weight(w, dw, torch.full_like(dw, 8.0), 16, c_begin=c0, c_end=c1, w4=False))
    out1 = torch.empty(B, N, H, H, device=cuda)
    ops.qgemm_i8_split(q, packs[0], packs[1], (d0, z0), (d1, z1), out1, H * H, bias=bias)
    out2 = torch.empty(B, N, H, H, device=cuda)
    ops.qgemm_i8(q, packs[0], d0, z0, out2, H * H, bias=bias)
    ops.qgemm_i8(q, packs[1], d1, z1, out2, H * H, a_c_offset=C0, accumulate=True)
    assert torch.equal(out1, out2)
    # each range on its own (validated against fp64 convolutions by the GEMM tests above), summed in fp64
    o0 = torch.empty(B, N, H, H, device=cuda)
    o1 = torch.empty(B, N, H, H, device=cuda)
    ops.qgemm_i8(q, packs[0], d0, z0, o0, H * H, bias=bias)
    ops.qgemm_i8(q, packs[1], d1, z1, o1, H * H, a_c_offset=C0)
    assert _rel_l2(out1.double(), o0.double() + o1.double()) < 1e-6


@pytest.mark.parametrize("zp", [128.0, 127.0, 0.0, 1.0])
def test_nhwc_producer_ties_round_to_even_quotient(cuda, zp):
    """x / delta exactly half-way between two integers must round to the EVEN quotient whatever the parity of the zero-point
    (torch.round semantics, quant_layer.py:267): the producers fold the zero-point into their magic-number rounding only when
    that keeps the tie rule"""
    from edadm import ops
    delta = 0.25
    k = torch.arange(-140, 140, dtype=torch.float32)
    vals = torch.cat([(k + 0.5) * delta, k * delta, (k + 0.25) * delta])              # ties, exact integers, plain values
    x = vals.repeat(64 * 16 * 16 * 4 // vals.numel() + 1)[:4 * 64 * 16 * 16].reshape(4, 64, 16, 16).to(cuda)
    d, z = torch.tensor([delta], device=cuda), torch.tensor([zp], device=cuda)
    q, _ = ops.act_quant_nhwc(x, ops.ActQuant(d, z, 256), 1)
    ref = torch.clamp(torch.round(x / d) + z, 0, 255).to(torch.uint8).permute(0, 2, 3, 1)
    assert torch.equal(q[:, 1:-1, 1:-1, :64], ref)


def test_conv_upsample_on_codes_is_exact(cuda):
    """`Upsample` with a conv (openaimodel.py Upsample): quantizing the low-resolution tensor and replicating the u8 codes equals
    quantizing the interpolated tensor, bit for bit, through the whole QuantModule"""
    from qdiff import QuantModel, set_weight_quantize_params, set_act_quantize_params
    from unet_zoo.ldm_unet import UNetModel, Upsample
    torch.manual_seed(23)
    model = UNetModel(image_size=16, in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1,
                      attention_resolutions=(), channel_mult=(1, 2), num_heads=1).to(cuda).eval()
    for p in model.parameters():
        if p.dim() > 1 and float(p.detach().abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
    wq = dict(n_bits=4, symmetric=True, channel_wise=True, scale_method='mse')
    aq = dict(n_bits=8, symmetric=True, channel_wise=False, scale_method='mse', leaf_param=True, prob=1.0)
    qnn = QuantModel(model, wq, aq, sm_abit=8).to(cuda).eval()
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(8, 4, 16, 16, generator=g).to(cuda)
    t = torch.randint(0, 1000, (8,), generator=g).to(cuda)
    with torch.no_grad():
        set_weight_quantize_params(qnn, (x, t))
        set_act_quantize_params(qnn, (x, t))
        qnn.set_quant_state(True, True)
        ups = [m for m in qnn.modules() if isinstance(m, Upsample) and m.use_conv]
        assert ups
        conv = ups[0].conv
        h = (torch.randn(8, conv.weight.shape[1], 8, 8, generator=g) * 0.7).to(cuda)
        y1 = conv.forward_upsample2x(h)
        assert conv.last_path == 'int8'
        y0 = conv(F.interpolate(h, scale_factor=2, mode="nearest"))
    assert y1.shape == y0.shape and torch.equal(y1, y0)


@pytest.mark.parametrize("channel_wise,bits,shape,shift", [(False, 8, (8, 64, 16, 16), 0.0), (False, 8, (4, 77, 320), 1.5), (True, 4, (96, 64, 3, 3), 0.0),
                                                            (True, 8, (40, 130), 0.0), (False, 8, (2, 8, 64, 64), 3.0)])
def test_scale_search_kernel_equals_tensor_op_search(cuda, channel_wise, bits, shape, shift):
    """edadm_mse_search_scores (all 100 candidates in one pass) lands on the same (delta, zero_point) as the candidate-by-candidate
    tensor-op search it replaces (reference quant_layer.py:150-213), incl. one-sided inputs (softmax-like, shift = 3)"""
    from qdiff.quant_layer import UniformAffineQuantizer, backend
    g = torch.Generator().manual_seed(bits * 7 + len(shape))
    x = (torch.randn(*shape, generator=g) * (0.1 if channel_wise else 1.3)).to(cuda)
    if shift == 3.0:
        x = torch.softmax(x * 3, -1)
    else:
        x = x + shift
    res = []
    for use_kernel in (True, False):
        q = UniformAffineQuantizer(n_bits=bits, symmetric=True, channel_wise=channel_wise, scale_method='mse', leaf_param=not channel_wise)
        backend.search_kernel = use_kernel
        try:
            q(x)
        finally:
            backend.search_kernel = True
        res.append((q.delta.detach().clone(), q.zero_point.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


@pytest.mark.parametrize("M,K,N,bias", [(4096, 384, 384, True), (8192, 384, 3072, False), (2048, 1536, 384, True), (300, 96, 100, True),
                                        (65536, 320, 64, False)])
def test_linear_bf16x3_forward_and_gradients(cuda, M, K, N, bias):
    """calibration-path GEMM (bf16 x 3 split on tcgen05, fp32 accumulation): forward, dgrad and wgrad (split-K + TMA reduce) against
    an fp64 reference -- far inside north_star's 1e-3, two orders of magnitude tighter than TF32"""
    from edadm import ops
    g = torch.Generator().manual_seed(M + N)
    x = (torch.randn(M, K, generator=g) * 1.7).to(cuda).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) * 0.08).to(cuda).requires_grad_(True)
    b = torch.randn(N, generator=g).to(cuda).requires_grad_(True) if bias else None
    gy = torch.randn(M, N, generator=g).to(cuda)
    y = ops.linear_bf16x3(x, w, b)
    y.backward(gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = F.linear(xd, wd, bd)
    yd.backward(gy.double())
    assert _rel_l2(y.detach().double(), yd.detach()) < 2e-5
    assert _rel_l2(x.grad.double(), xd.grad) < 2e-5
    assert _rel_l2(w.grad.double(), wd.grad) < 2e-5
    if bias:
        assert _rel_l2(b.grad.double(), bd.grad) < 1e-5
    # for scale: the same product in TF32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        e_tf32 = _rel_l2(F.linear(x.detach(), w.detach()).double(), F.linear(xd.detach(), wd.detach()))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    print(f"forward rel-L2 vs fp64: bf16x3 {_rel_l2((y.detach() - (b.detach() if bias else 0)).double(), F.linear(xd.detach(), wd.detach())):.1e}, TF32 {e_tf32:.1e}")


@pytest.mark.parametrize("B,C,N,H,W,R,bias", [(8, 192, 192, 32, 32, 3, True), (16, 384, 192, 16, 16, 3, True), (32, 576, 384, 8, 8, 3, False),
                                             (4, 128, 256, 32, 32, 1, True), (2, 64, 72, 64, 64, 3, True), (4, 200, 128, 16, 16, 3, False),
                                             (1, 32, 32, 128, 128, 3, True)])
def test_conv_bf16x3_forward_and_dgrad(cuda, B, C, N, H, W, R, bias):
    """calibration-path convolution (implicit GEMM on the bf16 x 3 kernel, NHWC split producer, NCHW TMA-store epilogue): forward,
    dgrad and wgrad (pixel-contiguous operands, tap shift by tensor-map coordinates, split-K reduce) against an fp64 convolution"""
    from edadm import ops
    g = torch.Generator().manual_seed(B * 131 + C)
    x = (torch.randn(B, C, H, W, generator=g) * 1.3).to(cuda).requires_grad_(True)
    w = (torch.randn(N, C, R, R, generator=g) * 0.05).to(cuda).requires_grad_(True)
    b = torch.randn(N, generator=g).to(cuda).requires_grad_(True) if bias else None
    kw = dict(stride=1, padding=(R - 1) // 2)
    assert ops.conv_bf16x3_ok(x, w, kw)
    gy = torch.randn(B, N, H, W, generator=g).to(cuda)
    y = ops.conv_bf16x3(x, w, b)
    y.backward(gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yd = F.conv2d(xd, wd, bd, **kw)
    yd.backward(gy.double())
    assert y.shape == yd.shape
    assert _rel_l2(y.detach().double(), yd.detach()) < 2e-5
    assert _rel_l2(x.grad.double(), xd.grad) < 2e-5
    assert _rel_l2(w.grad.double(), wd.grad) < 2e-5
    if bias:
        assert _rel_l2(b.grad.double(), bd.grad) < 1e-5


def test_conv_bf16x3_guard(cuda):
    from edadm import ops
    x = torch.randn(2, 64, 16, 16, device=cuda)
    assert not ops.conv_bf16x3_ok(x, torch.randn(64, 64, 3, 3, device=cuda), dict(stride=2, padding=1))
    assert not ops.conv_bf16x3_ok(x, torch.randn(4, 64, 3, 3, device=cuda), dict(stride=1, padding=1))
    assert not ops.conv_bf16x3_ok(torch.randn(2, 64, 12, 12, device=cuda), torch.randn(64, 64, 3, 3, device=cuda), dict(stride=1, padding=1))


@pytest.mark.parametrize("G,M,N,K", [(8, 256, 256, 384), (3, 1024, 1024, 96), (16, 128, 384, 256), (2, 256, 128, 1024)])
def test_bmm_nt_bf16x3_forward_and_gradients(cuda, G, M, N, K):
    """grouped calibration-path product C[g] = A[g] . B[g]^T (attention Q.K^T and P.V under autograd) vs fp64"""
    from edadm import ops
    g = torch.Generator().manual_seed(G * 7 + K)
    a = torch.randn(G, M, K, generator=g).to(cuda).requires_grad_(True)
    b = (torch.randn(G, N, K, generator=g) * 0.5).to(cuda).requires_grad_(True)
    gc = torch.randn(G, M, N, generator=g).to(cuda)
    assert ops.bmm_nt_bf16x3_ok(a, b)
    c = ops.bmm_nt_bf16x3(a, b)
    c.backward(gc)
    ad, bd = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    cd = torch.bmm(ad, bd.transpose(1, 2))
    cd.backward(gc.double())
    assert _rel_l2(c.detach().double(), cd.detach()) < 2e-5
    assert _rel_l2(a.grad.double(), ad.grad) < 2e-5
    assert _rel_l2(b.grad.double(), bd.grad) < 2e-5


# ------------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_optim(cuda):
    """edadm_fused_adam (both optimisers of a reconstruction unit in one pass over the flat gradient bucket) against the
    reference's optimiser, torch.optim.Adam (qdiff/block_recon.py:113-117), run in fp64 on the CPU: two parameter groups
    with their own (changing) learning rates, tensors whose sizes / flat offsets break 16-byte alignment, scalars (the
    activation step sizes), a tensor spanning several 8192-element segments; consumed gradients are cleared; version
    counters move so that packed-weight caches notice."""
    from qdiff.dist import GradBucket
    from qdiff._fused_adam import FusedAdam
    g = torch.Generator().manual_seed(71)
    shapes_w, shapes_a = [(40, 32, 3, 3), (7, 3), (20000,), (5, 5, 5)], [(), (), (1,), ()]
    init = [torch.randn(s, generator=g) for s in shapes_w] + [torch.rand(s, generator=g) * 0.1 + 0.01 for s in shapes_a]
    ours = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    ref = [torch.nn.Parameter(t.clone().double()) for t in init]
    nw = len(shapes_w)
    bucket = GradBucket(ours)
    lrs = torch.zeros(2, device=cuda)
    adam = FusedAdam(bucket, nw, lrs)
    opt_w, opt_a = torch.optim.Adam(ref[:nw], lr=1e-2, foreach=False), torch.optim.Adam(ref[nw:], lr=4e-4, foreach=False)
    versions = [p._version for p in ours]
    for it in range(6):
        lr_w, lr_a = 1e-2 * (1 - it / 8), 4e-4 * (1 - it / 8)
        lrs.copy_(torch.tensor([lr_w, lr_a]))
        for grp, lr in ((opt_w, lr_w), (opt_a, lr_a)):
            grp.param_groups[0]['lr'] = lr
        grads = [torch.randn(t.shape, generator=g) * (10.0 ** (-(i % 4))) for i, t in enumerate(init)]
        grads[2][:100] = 0.0                                   # exact zeros: update is 0 / (0 + eps)
        for p, r, gr in zip(ours, ref, grads):
            p.grad.copy_(gr.to(cuda))                          # views into bucket.flat
            r.grad = gr.double()
        adam.step()
        opt_w.step(); opt_a.step()
        assert float(bucket.flat.abs().max()) == 0.0           # consumed gradients cleared for the next backward
    adam.finish()
    assert all(p._version > v for p, v in zip(ours, versions))
    assert int(adam.step_count) == 6
    worst = 0.0
    for p, r, t0 in zip(ours, ref, init):
        moved = float((r.detach() - t0.double()).abs().max())
        err = float((p.detach().cpu().double() - r.detach()).abs().max())
        worst = max(worst, err / max(moved, 1e-30))
        assert err <= 1e-4 * moved + 1e-6 * float(t0.abs().max()), (tuple(t0.shape), err, moved)   # fp32 storage of p: ~1 ulp per step
    # moments against the fp64 optimiser state
    states = [opt.state[r] for r, opt in zip(ref, [opt_w] * nw + [opt_a] * (len(ref) - nw))]
    for key, buf in (('exp_avg', adam.exp_avg), ('exp_avg_sq', adam.exp_avg_sq)):
        want = torch.cat([st[key].reshape(-1) for st in states])
        got = buf.cpu().double()
        assert _rel_l2(got, want) < 1e-6
        off = 0
        for i, st in enumerate(states):         # per tensor, absolute: a scalar's first moment is a cancelling sum of its gradients
            n, scale = st[key].numel(), (10.0 ** (-(i % 4))) ** (1 if key == 'exp_avg' else 2)
            assert float((got[off:off + n] - want[off:off + n]).abs().max()) <= 1e-6 * scale, (key, i)
            off += n
    print(f"fused Adam vs fp64 torch.optim.Adam after 6 steps: worst |dp| / |total update| {worst:.2e}")
