"""CPU tests (-m "not gpu"): C-ABI library loads and exports every symbol include/edadm.h declares, host logic of the
qdiff drop-in (model rewrite, state switching, search == oracle, no-CPU-fallback behaviour), world_size-2 gloo tests of
the data-parallel plumbing."""
import os
import re
import subprocess
import sys

import pytest
import torch

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from edadm import native
    import importlib.util
    spec = importlib.util.spec_from_file_location("edadm_build", os.path.join(ROOT, "eda-dm_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    mod.build_lib()
    handle = native.load_library()
    header = open(os.path.join(ROOT, "include", "edadm.h")).read()
    declared = set(re.findall(r"\b(edadm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/edadm.h but not exported"
    assert declared == set(native.exported_symbols()), declared ^ set(native.exported_symbols())
    assert handle.edadm_abi_version() == 4
    assert handle.edadm_reduce_slots() > 0
    # the ctypes table must agree with the header on every prototype's parameter count (guards against ABI drift)
    protos = re.findall(r"\b(?:int|const char\*)\s+(edadm_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", re.sub(r"/\*.*?\*/", "", header, flags=re.S), flags=re.S)
    assert {n for n, _ in protos} == declared
    for name, args in protos:
        args = args.strip()
        n_args = 0 if args in ("", "void") else args.count(",") + 1
        assert n_args == len(native._SIGNATURES[name][1]), f"{name}: header has {n_args} parameters, ctypes table {len(native._SIGNATURES[name][1])}"


def test_argument_errors_are_reported_not_crashing():
    from edadm import native
    lib = native.lib
    with pytest.raises(native.EdadmError, match="null pointer"):
        lib.uaq_fwd(None, None, None, None, None, 16, 1, 1, 256, None, None, 1.0, 0, 0, None)


def test_no_cpu_fallback_for_quantized_path():
    from qdiff.quant_layer import UniformAffineQuantizer, QuantModule
    from edadm.native import EdadmError
    q = UniformAffineQuantizer(**H.AQ)
    with pytest.raises(EdadmError):
        q(torch.randn(4, 4))
    m = QuantModule(torch.nn.Linear(8, 8), H.WQ, H.AQ)
    assert m(torch.randn(2, 8)).shape == (2, 8)         # FP branch works on CPU
    m.set_quant_state(True, True)
    with pytest.raises(EdadmError):
        m(torch.randn(2, 8))


def _qmodel(model):
    from qdiff import QuantModel
    qnn = QuantModel(model, H.WQ, H.AQ, sm_abit=8)
    qnn.set_first_last_layer_to_8bit()
    qnn.disable_network_output_quantization()
    return qnn


def test_quant_model_rewrite_ddim():
    from qdiff.quant_layer import QuantModule, UniformAffineQuantizer
    from qdiff.quant_block import QuantResnetBlock, QuantAttnBlock, BaseQuantBlock
    g = H.load("ddim_tiny.npz")
    model = H.ddim_tiny_model()
    model.load_state_dict(H.state_dict(g))
    qnn = _qmodel(model)
    mods = dict(qnn.named_modules())
    assert isinstance(mods["model.down.0.block.0"], QuantResnetBlock)
    assert isinstance(mods["model.down.1.attn.0"], QuantAttnBlock)
    assert sum(isinstance(m, QuantModule) for m in mods.values()) == 51
    # first / last weight quantizers 8 bit (the "first" one is the time-embedding MLP)
    assert mods["model.temb.dense.0"].weight_quantizer.n_bits == 8
    assert mods["model.conv_out"].weight_quantizer.n_bits == 8
    assert mods["model.conv_in"].weight_quantizer.n_bits == 4
    assert mods["model.conv_out"].disable_act_quant is True
    # FP forward equals the reference's FP output, state switching toggles every unit
    x, t = torch.from_numpy(g["x"])[:4], torch.from_numpy(g["t"])[:4]
    with torch.no_grad():
        assert H.rel_l2(qnn(x, t), torch.from_numpy(g["y_fp"])) < 1e-6
    qnn.set_quant_state(True, False)
    assert all(m.use_weight_quant and not m.use_act_quant for m in mods.values() if isinstance(m, (QuantModule, BaseQuantBlock)))
    # golden quantizer names exist in the product model (after split twins are created by a split forward)
    qnn.set_quant_state(False, False)
    qnn.model.config.split_shortcut = True
    with torch.no_grad():
        qnn(x, t)
    mods = dict(qnn.named_modules())
    for name in H.qtable(g):
        assert isinstance(mods[name], UniformAffineQuantizer), name


@pytest.mark.parametrize("name", ["ldm_tiny.npz", "ldm_tiny_b.npz", "ldm_xattn_tiny.npz"])
def test_quant_model_rewrite_ldm(name):
    from qdiff.quant_layer import UniformAffineQuantizer
    from qdiff.quant_block import QuantResBlock, QuantQKMatMul, QuantSMVMatMul, QuantBasicTransformerBlock
    g = H.load(name)
    model = H.ldm_model(name)
    model.load_state_dict(H.state_dict(g))
    qnn = _qmodel(model)
    kinds = {type(m) for m in qnn.modules()}
    assert QuantResBlock in kinds
    assert (QuantBasicTransformerBlock in kinds) == ("xattn" in name)
    if "xattn" not in name:
        assert QuantQKMatMul in kinds and QuantSMVMatMul in kinds
    args = [torch.from_numpy(g["x"])[:4], torch.from_numpy(g["t"])[:4]] + ([torch.from_numpy(g["ctx"])[:4]] if "ctx" in g.files else [])
    qnn.model.split_shortcut = True
    with torch.no_grad():
        assert H.rel_l2(qnn(*args), torch.from_numpy(g["y_fp"])) < 1e-6
    mods = dict(qnn.named_modules())
    for qname in H.qtable(g):
        assert isinstance(mods[qname], UniformAffineQuantizer), qname


def test_lazy_concatenation_host_logic():
    """CatPair quacks like the concatenated tensor and materialises on demand; a QuantResBlock only takes the two-source route on
    the integer path (CUDA, no hooks, no gradients): on CPU / in FP state the UNet forward falls back to th.cat and stays exact"""
    from edadm import ops
    from qdiff.quant_block import QuantResBlock
    a, b = torch.randn(2, 32, 4, 4), torch.randn(2, 16, 4, 4)
    pair = ops.CatPair(a, b)
    assert tuple(pair.shape) == (2, 48, 4, 4) and pair.dim() == 4 and pair.numel() == a.numel() + b.numel()
    assert pair.dtype == a.dtype and pair.device == a.device and not pair.is_cuda
    full = pair.materialize()
    assert torch.equal(full, torch.cat([a, b], 1)) and pair.materialize() is full          # cached
    assert not ops.cat_slices_ok(pair, 0)                                                   # CPU tensors never take the slice kernels
    g = H.load("ldm_tiny.npz")
    model = H.ldm_model("ldm_tiny.npz")
    model.load_state_dict(H.state_dict(g))
    qnn = _qmodel(model)
    blocks = [m for m in qnn.modules() if isinstance(m, QuantResBlock)]
    assert blocks and all(m.lazy_cat(a, b) is None for m in blocks)                         # CPU: always the materialised path
    args = [torch.from_numpy(g["x"])[:4], torch.from_numpy(g["t"])[:4]]
    with torch.no_grad():
        assert H.rel_l2(qnn(*args), torch.from_numpy(g["y_fp"])) < 1e-6


def test_scale_search_equals_oracle_on_cpu():
    """The product's range search (pure tensor ops, device independent) == the oracle's restatement == the reference."""
    from qdiff.quant_layer import UniformAffineQuantizer
    u = H.load("unit.npz")
    q = UniformAffineQuantizer(**H.AQ)
    for key in ("act_x0", "act_x1"):
        d, z = q.init_quantization_scale_1(torch.from_numpy(u[key]), False)
    assert torch.equal(d, torch.from_numpy(u["act_delta"])) and torch.equal(z, torch.from_numpy(u["act_zp"]))
    for bits in (4, 8):
        p = dict(H.WQ); p["n_bits"] = bits
        q = UniformAffineQuantizer(**p)
        d, z = q.init_quantization_scale_1(torch.from_numpy(u["w"]), True)
        assert torch.equal(d, torch.from_numpy(u[f"w{bits}_delta"])) and torch.equal(z, torch.from_numpy(u[f"w{bits}_zp"]))
    # asymmetric two-sided input -> 2-D search; compare with the reference-style brute force on a tiny tensor
    p = dict(H.AQ); p.update(symmetric=False, n_bits=3, leaf_param=False)
    q = UniformAffineQuantizer(**p)
    x = torch.tensor([-0.3, -0.1, 0.0, 0.2, 0.5, 0.9, 1.4])
    d, z = q.init_quantization_scale_1(x, False)
    assert 0.0 < float(d) < 1.0 and 0.0 <= float(z) <= 7.0


def test_implicit_tiling_rule_matches_c_side():
    from qdiff.quant_layer import _implicit_tiling_ok
    assert _implicit_tiling_ok(100, 32, 32) and _implicit_tiling_ok(3, 64, 64) and _implicit_tiling_ok(8, 4, 4)
    assert _implicit_tiling_ok(1, 1, 1024) and _implicit_tiling_ok(5, 2, 2)
    assert not _implicit_tiling_ok(2, 12, 12) and not _implicit_tiling_ok(2, 17, 17)
    assert not _implicit_tiling_ok(1, 3, 64)            # 2 rows per tile do not divide 3


def test_walker_visits_units_in_reference_order():
    from qdiff._walker import UnitWalker
    g = H.load("ddim_tiny.npz")
    qnn = _qmodel(H.ddim_tiny_model())
    order = []
    names = {id(m): n for n, m in qnn.named_modules()}
    UnitWalker(lambda m: order.append(("layer", names[id(m)])), lambda m: order.append(("block", names[id(m)]))).walk(qnn)
    flat = [n for _, n in order]
    assert flat[0] == "model.temb.dense.0" and flat[1] == "model.temb.dense.1" and flat[2] == "model.conv_in"
    assert flat[3] == "model.down.0.block.0"
    # `up` is walked from its last level to its first; level 1 (attention) unrolled block/attn pairs then upsample
    up = [n for n in flat if ".up." in n]
    assert up[0].startswith("model.up.1.block.0") and up[1].startswith("model.up.1.attn.0")
    assert up[-1].startswith("model.up.0.block")
    assert flat[-1] == "model.conv_out"
    kinds = dict((n, k) for k, n in order)
    assert kinds["model.down.0.downsample.conv"] == "layer"


def test_fused_adam_segment_table():
    """host side of edadm_fused_adam: every parameter element belongs to exactly one segment, segments never straddle two
    tensors, flat offsets follow the GradBucket order, the group bit separates alphas from step sizes"""
    from qdiff._fused_adam import segment_table, SEGMENT
    params = [torch.zeros(3, 5), torch.zeros(2 * SEGMENT + 17), torch.zeros(()), torch.zeros(SEGMENT), torch.zeros(1)]
    table, total = segment_table(params, n_group0=2)
    assert total == sum(p.numel() for p in params) and table.dtype.name == "int64" and table.shape[1] == 3
    flat_cover, off = [], 0
    for i, p in enumerate(params):
        rows = [r for r in table if p.data_ptr() <= r[0] < p.data_ptr() + 4 * max(1, p.numel())]
        assert sum(int(r[2]) & 0xFFFFFFFF for r in rows) == p.numel()
        for r in rows:
            cnt, grp = int(r[2]) & 0xFFFFFFFF, int(r[2]) >> 32
            assert 1 <= cnt <= SEGMENT and grp == (0 if i < 2 else 1)
            assert (int(r[0]) - p.data_ptr()) // 4 == int(r[1]) - off          # same element in the tensor and in the flat buffer
            flat_cover.append((int(r[1]), cnt))
        off += p.numel()
    flat_cover.sort()
    assert flat_cover[0][0] == 0 and all(a + n == b for (a, n), (b, _) in zip(flat_cover, flat_cover[1:]))
    assert flat_cover[-1][0] + flat_cover[-1][1] == total


def test_device_learning_rate_table_equals_torch_scheduler():
    """the captured reconstruction step reads its learning rates from a table indexed by a device counter (`_cosine_lr`); the
    reference steps torch's CosineAnnealingLR(T_max=iters, eta_min=0) (block_recon.py:114-117, 203-206): same schedule"""
    from qdiff._recon_engine import _cosine_lr, LinearTempDecay
    for lr0, iters in ((1e-2, 40), (4e-4, 1000), (1e-2, 7)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adam([p], lr=lr0)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=iters, eta_min=0.)
        for t in range(iters):
            lr = opt.param_groups[0]['lr']
            assert abs(lr - _cosine_lr(lr0, t, iters)) <= 1e-12 * lr0, (lr0, iters, t)
            p.grad = torch.ones(1)
            opt.step()
            sched.step()
    # temperature of the (disabled by every caller) rounding regulariser, block_recon.py:305-323: linear from start_b to end_b
    td = LinearTempDecay(100, rel_start_decay=0.2, start_b=20, end_b=2)
    assert td(1) == 20 and td(19) == 20 and td(20) == 20 and abs(td(60) - 11.0) < 1e-12 and td(100) == 2 and td(150) == 2


def test_grad_bucket_views_and_sharding():
    from qdiff import dist as qdist
    a = torch.nn.Parameter(torch.zeros(3, 4))
    b = torch.nn.Parameter(torch.zeros(()))
    bucket = qdist.GradBucket([a, b])
    (a.sum() * 2 + b * 3).backward()
    assert bucket.flat.tolist() == [2.0] * 12 + [3.0]
    assert a.grad.data_ptr() == bucket.flat.data_ptr()
    bucket.zero()
    assert float(a.grad.abs().sum()) == 0.0
    assert [qdist.shard_rows(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


DIST_SCRIPT = r'''
import os, sys
sys.path[:0] = [os.environ["EDADM_ROOT"], os.path.join(os.environ["EDADM_ROOT"], "eda-dm_b200")]
import torch, torch.distributed as dist
from qdiff import dist as qdist
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
assert qdist.is_active() and qdist.world_size() == 2
x = torch.arange(20.).reshape(10, 2)
(shard,) = qdist.shard_calibration([x])
assert shard.shape[0] == 5 and float(shard[0, 0]) == (0.0 if r == 0 else 10.0)
# identical parameters on both ranks, rank-dependent gradients -> mean after the single flat all-reduce
p1 = torch.nn.Parameter(torch.ones(4)); p2 = torch.nn.Parameter(torch.ones(()))
bucket = qdist.GradBucket([p1, p2])
((p1 * (r + 1)).sum() + p2 * (10 * (r + 1))).backward()
bucket.all_reduce_mean()
assert torch.allclose(p1.grad, torch.full((4,), 1.5)) and abs(float(p2.grad) - 15.0) < 1e-6
# data-parallel equivalence: mean over ranks of per-shard mean-loss gradients == gradient of the global mean loss
g = torch.Generator().manual_seed(0)
data = torch.randn(8, 3, generator=g); tgt = torch.randn(8, generator=g)
wt = torch.nn.Parameter(torch.zeros(3))
bucket = qdist.GradBucket([wt])
lo, hi = qdist.shard_rows(8)
((data[lo:hi] @ wt - tgt[lo:hi]) ** 2).mean().backward()
bucket.all_reduce_mean()
wref = torch.zeros(3, requires_grad=True)
((data @ wref - tgt) ** 2).mean().backward()
assert torch.allclose(wt.grad, wref.grad, atol=1e-6)
opt = torch.optim.Adam([wt], lr=0.1); opt.step()
gathered = [torch.zeros(3) for _ in range(w)]
dist.all_gather(gathered, wt.detach())
assert torch.equal(gathered[0], gathered[1])           # replicas stay in lock-step without a broadcast
dist.barrier(); dist.destroy_process_group()
print("rank", r, "ok")
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dist_check.py"
    script.write_text(DIST_SCRIPT)
    env = dict(os.environ, EDADM_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_staged_cache_equals_prefix_rerun_fp_path():
    """f2: the staged cache builder (frontier state reuse) returns the tensors the reference-style builder computes by re-running
    the prefix -- FP path on the CPU here (the quantized path is covered by the -m gpu twin), units visited in walk order, plus a
    backwards request (forces a rebuild) and the staged == plain forward identity of the zoo UNets"""
    import torch
    from unet_zoo.ddpm_unet import DDPMUNet
    from unet_zoo.ldm_unet import UNetModel
    from qdiff import QuantModel
    from qdiff.quant_layer import backend, QuantModule
    from qdiff.quant_block import BaseQuantBlock
    from qdiff.data_utils import save_inp_oup_data
    wq = {'n_bits': 4, 'symmetric': True, 'channel_wise': True, 'scale_method': 'mse'}
    aq = {'n_bits': 8, 'symmetric': True, 'channel_wise': False, 'scale_method': 'mse', 'leaf_param': True, 'prob': 1.0}
    torch.manual_seed(0)
    cases = [(DDPMUNet(ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), resolution=16, dropout=0.0), (3, 16, 16), None),
             (UNetModel(image_size=8, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[1, 2], num_res_blocks=1,
                        channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True, transformer_depth=1, context_dim=24), (3, 8, 8), (3, 24))]
    for fp, shape, ctx in cases:
        fp = fp.eval()
        g = torch.Generator().manual_seed(1)
        cali = [torch.randn(8, *shape, generator=g), torch.randint(0, 1000, (8,), generator=g)]
        if ctx:
            cali.append(torch.randn(8, *ctx, generator=g))
        with torch.no_grad():
            y_plain = fp(*cali)
        qnn = QuantModel(fp, wq, aq, sm_abit=8).eval()
        with torch.no_grad():
            assert torch.equal(qnn(*cali), y_plain)            # staged forward of the rewritten model == the FP model's output
        units = [m for m in qnn.model.modules() if isinstance(m, BaseQuantBlock) and type(m).__name__ not in ("QuantQKMatMul", "QuantSMVMatMul")]
        units += [m for m in qnn.model.modules() if isinstance(m, QuantModule)][:3]
        order = sorted(units, key=lambda m: [id(x) for x in qnn.model.modules()].index(id(m)))
        for unit in order + [order[1]]:
            got = []
            for reuse in (True, False):
                backend.cache_prefix_reuse = reuse
                try:
                    got.append(save_inp_oup_data(qnn, unit, cali, asym=False, act_quant=False, batch_size=4, input_prob=True, keep_gpu=True))
                finally:
                    backend.cache_prefix_reuse = True
            (r0, i0, o0), (r1, i1, o1) = got
            assert r0 == r1 and torch.equal(o0, o1)
            flat = lambda t: [a for pair in t for a in (pair if isinstance(pair, (list, tuple)) else [pair])]
            for a, b in zip(flat(i0), flat(i1)):
                assert torch.equal(a, b)


def test_tdac_allocation_matches_reference_loops():
    """f4: Gram-matrix TDAC scores / allocation / assembly == the reference's O(T^2) loops (oracle/tdac_oracle.py)"""
    import torch
    from qdiff import tdac
    from oracle import tdac_oracle
    g = torch.Generator().manual_seed(3)
    T_, N = 12, 6
    base = torch.randn(N, 16, 4, 4, generator=g)
    feats = [base * (1.0 + 0.35 * k) + 0.8 * torch.randn(N, 16, 4, 4, generator=g) * (k % 3) for k in range(T_)]
    d0, c0 = tdac_oracle.scores(feats, dense_r=3.0)
    d1, c1 = tdac.tdac_scores(feats, dense_r=3.0)
    assert torch.equal(d0, d1) and 0 < int(d0.max()) and int(d0.min()) < int(d0.max())
    assert torch.allclose(c0, c1, rtol=1e-4, atol=1e-3)
    for lam in (0.0, 1.0, 2.5):
        assert torch.equal(tdac_oracle.allocation(feats, lam, 64), tdac.tdac_allocation(feats, lam, 64))
    t_num = tdac.tdac_allocation(feats, 1.0, 24)
    traj = [torch.randn(N, 3, 8, 8, generator=g) for _ in range(T_)]
    calib, t, ts = tdac.tdac_assemble(traj, t_num, N, seq=range(0, 1000, 1000 // T_)[:T_], generator=torch.Generator().manual_seed(9))
    assert torch.equal(calib, tdac_oracle.assemble(traj, t, N)) and calib.shape[0] == 24
    assert torch.equal(torch.bincount(t, minlength=T_), t_num)
    assert ts.shape == t.shape
