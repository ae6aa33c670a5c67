"""QuantModule / UniformAffineQuantizer with the reference's interface (qdiff/quant_layer.py:36-446 of
BienLuky/EDA-DM) on top of the sm_100a kernels in libedadm.so.

Two execution paths, chosen per call by `QuantModule.forward`:

* integer path (sampling, cache building: weight+act quantization on, no gradient wanted) --
  activations are quantized to u8 codes, weights are packed once to s8 codes, and the conv / linear
  runs as an exact int32 tcgen05 GEMM with the dequant in the epilogue (edadm_qgemm_i8);
* calibration path (reconstruction, weight-only quantization, scale search) -- the fused fake-quant
  kernels (edadm_uaq_fwd/bwd, edadm_adaround_fwd/bwd) produce fp32 operands with straight-through
  gradients, feeding a library fp32 convolution.

There is no CPU fallback: CPU tensors are only accepted on the un-quantized (FP) branch.
"""
import logging
from typing import Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from edadm import ops
from edadm.native import EdadmError

logger = logging.getLogger(__name__)


class _Backend:
    """Process-wide knobs of the calibration path (defaults reproduce the reference's fp32 numerics)."""
    allow_tf32 = False       # library conv/matmul of the calibration path in strict fp32
    integer_path = True      # use the tcgen05 int8 GEMM whenever it applies
    fuse_norm = True         # GroupNorm + SiLU + activation quantizer as one producer pass on the integer path
    fast_silu = False        # True: SiLU in the fused GroupNorm producer through the SFU ex2 / rcp approximations (~3 ulp, a few 1e-4 of
                             # the codes move by one step); False: ATen's exact forms, codes equal the module-by-module CUDA path
    fused_attention = True   # quantized attention as one tcgen05 kernel (edadm_qattn_fwd); False: fake-quant kernels around library bmm
    search_kernel = True     # scale search: all 100 clipping candidates scored in one pass (edadm_mse_search_scores); False: tensor ops
    lazy_cat = True          # the up path's skip concatenation is never written: its consumers read the two sources in place (QuantResBlock.lazy_cat)
    fuse_epilogue = True     # linears whose only consumer is the next activation quantizer emit its u8 codes from the GEMM epilogue
    # (4-bit weight storage with in-smem unpack: `edadm.ops.w4_storage`)
    recon_cuda_graph = True  # capture the reconstruction iteration in one CUDA graph after 3 eager iterations
    recon_graph_checkpointed = True   # ... including units whose backward recomputes the forward (transformer / attention blocks)
    recon_overlap_allreduce = True    # data parallel: all-reduce each alpha gradient as soon as its backward has produced it
    recon_memoise_fp_taps = True      # FP-model taps of the per-layer loss: computed once per unit for all cached samples (HBM), not per iteration
    recon_memoise_bytes = 32 << 30    #   ... as long as they fit in this many bytes
    recon_fused_adam = True           # both Adam updates of an iteration as one pass over the flat gradient bucket (edadm_fused_adam)
    calib_gemm_bf16x3 = True          # linears of the reconstruction loop (fwd / dgrad / wgrad) on edadm_gemm_bf16x3 instead of cuBLAS fp32
    calib_conv_wgrad_bf16x3 = True    # ... including the convolution wgrad (False: cuDNN's, TF32 by torch's default -- 6-10 % faster on the
                                      # 576-channel ImageNet ResBlocks, 10-18 % slower on church's, 50x less accurate)
    in_recon = False                  # set by the reconstruction engine around its loop (FP-target forwards included)
    cache_prefix_reuse = True         # calibration cache builder keeps the network state at the frontier of the finished units (f2)
    recon_overlap_fp = False  # ... with the FP forward on a forked stream (a parallel graph branch): +4 % on a church
                              # 16x16 ResBlock, -23 % on an ImageNet 32x32 one (measured), hence opt-in
    qdrop_inkernel_rng = False  # False: QDrop masks come from torch.rand_like (the reference's stream, graph-safe);
                                # True: drawn inside the kernel (Philox4x32, no extra memory pass)
    code_tap = None          # test hook: callable(module, codes, pad) that sees the u8 activation codes the integer path consumes
    qdrop_seed = None        # None -> torch.initial_seed()
    qdrop_offset = 0         # running Philox offset (one fresh sub-stream per fake-quant call)

    @classmethod
    def next_stream(cls, numel: int):
        seed = cls.qdrop_seed if cls.qdrop_seed is not None else torch.initial_seed()
        off = cls.qdrop_offset
        cls.qdrop_offset += (numel + 3) // 4 * 4
        return seed & 0xFFFFFFFFFFFFFFFF, off


backend = _Backend


class StraightThrough(nn.Module):
    def __init__(self, channel_num: int = 1):
        super().__init__()

    def forward(self, input):
        return input


def _tdiv(t: torch.Tensor, scalar: float) -> torch.Tensor:
    """IEEE division of a tensor by a Python scalar.  torch's CUDA kernel turns `t / python_scalar` into
    `t * (1/scalar)`, which is 1 ulp off and flips `round(min/scale)` between 7 and 8 (or 127 and 128) during the range
    search; dividing by a 0-dim DEVICE tensor keeps true division, i.e. the CPU semantics the oracle is pinned to."""
    return t / torch.full((), float(scalar), dtype=t.dtype, device=t.device)


def round_ste(x: torch.Tensor):
    """Round with identity gradient (reference quant_layer.py:19-23)."""
    return (x.round() - x).detach() + x


def lp_loss(pred, tgt, p=2.0, reduction='none'):
    """L_p loss (reference quant_layer.py:26-33); one fused reduction kernel on CUDA tensors."""
    if pred.is_cuda:
        return ops.lp_loss(pred, tgt, p, reduction)
    d = (pred - tgt).abs().pow(p)
    return d.sum(1).mean() if reduction == 'none' else d.mean()


class UniformAffineQuantizer(nn.Module):
    """Uniform affine fake-quantizer; same constructor, attributes and init behaviour as the reference
    (quant_layer.py:36-358): `delta`/`zero_point` are found by an L2.4 grid search on the first
    un-inited forward, `delta` becomes an nn.Parameter when `leaf_param`, and `inited` only changes
    through `set_inited()`.
    """

    def __init__(self, n_bits: int = 8, symmetric: bool = False, channel_wise: bool = False, scale_method: str = 'max',
                 leaf_param: bool = False, always_zero: bool = False, prob: float = 1.0):
        super().__init__()
        self.sym = symmetric
        self.bitwidth_refactor(n_bits)
        self.delta = None
        self.zero_point = None
        self.inited = False
        self.leaf_param = leaf_param
        self.channel_wise = channel_wise
        self.scale_method = scale_method
        self.running_stat = False
        self.always_zero = always_zero
        if self.leaf_param:
            self.x_min, self.x_max = None, None
        self.running_min = None
        self.running_max = None
        self.one_side_dist = None
        self.num = 100
        self.eps = torch.tensor(1e-8, dtype=torch.float32)
        self.prob = prob
        self.is_training = False

    def set_inited(self, inited: bool = True):
        self.inited = inited

    def bitwidth_refactor(self, refactored_bit: int):
        self.n_bits = refactored_bit
        self.n_levels = 2 ** self.n_bits

    # ---- scale search (reference quant_layer.py:79-244); init-time, plain tensor ops on the device ----
    def update_quantize_range(self, x_min, x_max):
        if self.running_min is None:
            self.running_min, self.running_max = x_min, x_max
        self.running_min = 0.1 * x_min + 0.9 * self.running_min
        self.running_max = 0.1 * x_max + 0.9 * self.running_max
        return self.running_min, self.running_max

    def calculate_qparams(self, min_val, max_val):
        qmax = self.n_levels - 1
        lo = torch.clamp(min_val, max=0.0)
        hi = torch.clamp(max_val, min=0.0)
        scale = torch.clamp(_tdiv(hi - lo, qmax), min=1e-8)
        zero_point = torch.clamp(0 - torch.round(lo / scale), 0, qmax)
        return scale, zero_point

    def _score_candidates(self, x_flat, new_min, new_max, rows: bool):
        """L2.4 score of each (new_min, new_max) candidate; `rows` = per-channel candidates [C] against
        x_flat [C, n], else a batch of per-tensor candidates [k] against x_flat [1, n]."""
        scale, zp = self.calculate_qparams(new_min, new_max)
        scale, zp = scale.reshape(-1, 1), zp.reshape(-1, 1)
        q = torch.clamp(torch.round(x_flat / scale) + zp, 0, self.n_levels - 1)
        return ((q - zp) * scale - x_flat).abs_().pow_(2.4).mean(1)

    def perform_1D_search(self, x):
        if self.channel_wise:
            y = torch.flatten(x, 1)
            x_min, x_max = y.amin(1), y.amax(1)
        else:
            y = x.reshape(1, -1)
            x_min, x_max = x.amin(), x.amax()
        xrange = torch.max(x_min.abs(), x_max)
        steps = torch.arange(1, self.num + 1, device=x.device)
        if backend.search_kernel and x.is_cuda and x.dtype == torch.float32 and self.num <= 128 and y.shape[0] <= 65535:
            # every candidate of every channel scored in ONE pass over the tensor (edadm_mse_search_scores); the candidate
            # (delta, zero_point) pairs are the same fp32 values the loop below would try, in the same order
            thres = _tdiv(xrange, self.num).reshape(-1, 1) * steps.reshape(1, -1).to(x.dtype)           # [S, num]
            new_min = torch.zeros_like(thres) if self.one_side_dist == 'pos' else -thres
            new_max = torch.zeros_like(thres) if self.one_side_dist == 'neg' else thres
            scale, zp = self.calculate_qparams(new_min, new_max)
            scores = ops.mse_search_scores(y, scale, zp, self.n_levels, 2.4)
            ind = torch.argmin(scores, dim=1, keepdim=True)        # first minimum, like the strict `<` of the reference loop
            best_min, best_max = new_min.gather(1, ind).reshape(-1), new_max.gather(1, ind).reshape(-1)
            if self.channel_wise:
                # candidates never beat the reference's initial best_score of 1e10 only if every score is >= 1e10
                return best_min, best_max
            return best_min[0], best_max[0]
        if not self.channel_wise:
            thres = _tdiv(xrange, self.num) * steps
            new_min = torch.zeros_like(thres) if self.one_side_dist == 'pos' else -thres
            new_max = torch.zeros_like(thres) if self.one_side_dist == 'neg' else thres
            scores = [self._score_candidates(y, new_min[i:i + 8], new_max[i:i + 8], False) for i in range(0, self.num, 8)]
            ind = torch.argmin(torch.cat(scores))
            return new_min[ind], new_max[ind]
        best_score = torch.full_like(x_min, 1e10)
        best_min, best_max = x_min.clone(), x_max.clone()
        for i in range(1, self.num + 1):
            thres = _tdiv(xrange, self.num) * i
            new_min = torch.zeros_like(x_min) if self.one_side_dist == 'pos' else -thres
            new_max = torch.zeros_like(x_max) if self.one_side_dist == 'neg' else thres
            score = self._score_candidates(y, new_min, new_max, True)
            better = score < best_score
            best_min = torch.where(better, new_min, best_min)
            best_max = torch.where(better, new_max, best_max)
            best_score = torch.min(score, best_score)
        return best_min, best_max

    def perform_2D_search(self, x):
        if self.channel_wise:
            y = torch.flatten(x, 1)
            x_min, x_max = torch.clamp(y.amin(1), max=0.0), torch.clamp(y.amax(1), min=0.0)
        else:
            y = x.reshape(1, -1)
            x_min, x_max = x.amin().reshape(1), x.amax().reshape(1)
        xrange = x_max - x_min
        best_score = torch.full_like(x_min, 1e10)
        best_min, best_max = x_min.clone(), x_max.clone()
        for i in range(1, self.num + 1):
            tmp_max = _tdiv(xrange, self.num) * i
            tmp_delta = _tdiv(tmp_max, 2 ** self.n_bits - 1)
            for zp in range(0, self.n_levels):
                new_min, new_max = -zp * tmp_delta, tmp_max - zp * tmp_delta
                score = self._score_candidates(y, new_min, new_max, True)
                better = score < best_score
                best_min = torch.where(better, new_min, best_min)
                best_max = torch.where(better, new_max, best_max)
                best_score = torch.min(best_score, score)
        if not self.channel_wise:
            return best_min[0], best_max[0]
        return best_min, best_max

    def get_x_min_x_max(self, x):
        if self.scale_method != 'mse':
            raise NotImplementedError
        if self.one_side_dist is None:
            self.one_side_dist = 'pos' if x.min() >= 0.0 else 'neg' if x.max() <= 0.0 else 'no'
        if self.one_side_dist != 'no' or self.sym:
            best_min, best_max = self.perform_1D_search(x)
        else:
            best_min, best_max = self.perform_2D_search(x)
        if self.leaf_param:
            return self.update_quantize_range(best_min, best_max)
        return best_min, best_max

    def init_quantization_scale_1(self, x: torch.Tensor, channel_wise: bool = False):
        with torch.no_grad():
            x_min, x_max = self.get_x_min_x_max(x.detach())
            delta, zero_point = self.calculate_qparams(x_min, x_max)
        if channel_wise:
            shape = [1] * x.dim()
            shape[0] = x.shape[0]
            delta, zero_point = delta.reshape(shape), zero_point.reshape(shape)
        return delta, zero_point

    def init_quantization_scale_2(self, x: torch.Tensor, channel_wise: bool = False):
        """'max' / 'max_scale' range (reference quant_layer.py:278-345): the tensor's (per-channel) extrema instead of a
        search.  The reference walks the channels in a Python loop with `.item()` on every one; here the same float64 host
        arithmetic (`x_absmax / n_levels`, `round(-x_min / delta)`) is applied to the vector of channel extrema."""
        if 'max' not in self.scale_method:
            raise NotImplementedError
        with torch.no_grad():
            y = torch.flatten(x.detach(), 1) if channel_wise else x.detach().reshape(1, -1)
            lo64, hi64 = y.amin(1).double(), y.amax(1).double()
            if self.leaf_param and not channel_wise:
                self.x_min, self.x_max = x.data.min(), x.data.max()
            x_min, x_max = lo64.clamp(max=0.0), hi64.clamp(min=0.0)
            if 'scale' in self.scale_method:
                x_min, x_max = x_min * (self.n_bits + 2) / 8, x_max * (self.n_bits + 2) / 8
            if self.sym:
                delta = torch.max(x_min.abs(), x_max) / self.n_levels
            else:
                delta = (hi64 - lo64) / (self.n_levels - 1)
            delta = torch.where(delta < 1e-8, torch.full_like(delta, 1e-8), delta)
            if self.sym or self.always_zero:
                zero_point = torch.zeros_like(delta)
            else:
                zero_point = torch.round(-x_min / delta)
            delta, zero_point = delta.to(x.dtype), zero_point.to(x.dtype)
        if channel_wise:
            shape = [1] * x.dim()
            shape[0] = x.shape[0]
            return delta.reshape(shape), zero_point.reshape(shape)
        return delta.reshape(()), zero_point.reshape(())

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        if self.inited is False:
            if self.scale_method == 'mse':
                delta, self.zero_point = self.init_quantization_scale_1(x, self.channel_wise)
            elif self.scale_method == 'max':
                delta, self.zero_point = self.init_quantization_scale_2(x, self.channel_wise)
            else:
                raise NotImplementedError
            self.delta = torch.nn.Parameter(delta) if self.leaf_param else delta
        if not x.is_cuda:
            raise EdadmError("UniformAffineQuantizer needs CUDA tensors: the fake-quant kernel has no CPU fallback")
        if self.is_training and self.prob < 1.0:
            if backend.qdrop_inkernel_rng:
                seed, offset = backend.next_stream(x.numel())
                return ops.uaq_fake_quant(x, self.delta, self.zero_point, self.n_levels, None, self.prob, seed, offset)
            # torch.where(torch.rand_like(x) < prob, x_dequant, x) -- reference quant_layer.py:271-272, same draws
            return ops.uaq_fake_quant(x, self.delta, self.zero_point, self.n_levels, None, self.prob, 0, 0,
                                      keep_rand=torch.rand_like(x))
        return ops.uaq_fake_quant(x, self.delta, self.zero_point, self.n_levels)

    def codes(self, x: torch.Tensor):
        """Integer codes clamp(round(x/delta)+zp, 0, L-1) as uint8 (never materialised by the reference)."""
        return ops.uaq_forward(x, self.delta, self.zero_point, self.n_levels, want_codes=True)[1]

    def extra_repr(self):
        s = 'bit={n_bits}, scale_method={scale_method}, symmetric={sym}, channel_wise={channel_wise},' \
            ' leaf_param={leaf_param}'
        return s.format(**self.__dict__)


def _version_of(t):
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape))


class QuantModule(nn.Module):
    """Quantized Conv2d / Conv1d / Linear with the reference's constructor, attributes and
    `forward(input, split=0)` (quant_layer.py:360-446)."""

    def __init__(self, org_module: Union[nn.Conv2d, nn.Linear, nn.Conv1d], weight_quant_params: dict = {},
                 act_quant_params: dict = {}, disable_act_quant: bool = False, act_quant_mode: str = 'qdiff'):
        super().__init__()
        self.weight_quant_params = weight_quant_params
        self.act_quant_params = act_quant_params
        if isinstance(org_module, nn.Conv2d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv2d
        elif isinstance(org_module, nn.Conv1d):
            self.fwd_kwargs = dict(stride=org_module.stride, padding=org_module.padding,
                                   dilation=org_module.dilation, groups=org_module.groups)
            self.fwd_func = F.conv1d
        else:
            self.fwd_kwargs = dict()
            self.fwd_func = F.linear
        self.weight = org_module.weight
        self.org_weight = org_module.weight.data.clone()
        if org_module.bias is not None:
            self.bias = org_module.bias
            self.org_bias = org_module.bias.data.clone()
        else:
            self.bias = None
            self.org_bias = None
        self.use_weight_quant = False
        self.use_act_quant = False
        self.act_quant_mode = act_quant_mode
        self.disable_act_quant = disable_act_quant
        self.weight_quantizer = UniformAffineQuantizer(**self.weight_quant_params)
        if self.act_quant_mode == 'qdiff':
            self.act_quantizer = UniformAffineQuantizer(**self.act_quant_params)
        self.split = 0
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False
        self.extra_repr = org_module.extra_repr
        self._packed = None       # (key, [PackedWeight, ...]) cache of the integer path
        self.last_path = None     # 'fp' | 'int8' | 'fake' -- which branch the last forward took

    # non-persistent caches must not leak into copies / state_dict
    def _apply(self, fn, *a, **k):
        self._packed = None
        out = super()._apply(fn, *a, **k)
        self.org_weight = fn(self.org_weight)
        if self.org_bias is not None:
            self.org_bias = fn(self.org_bias)
        return out

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant = weight_quant
        self.use_act_quant = act_quant

    def set_split(self):
        self.weight_quantizer_0 = UniformAffineQuantizer(**self.weight_quant_params)
        if self.act_quant_mode == 'qdiff':
            self.act_quantizer_0 = UniformAffineQuantizer(**self.act_quant_params)

    # ---- path selection ------------------------------------------------------------------------------
    def _quantizers(self):
        if self.split != 0:
            return [self.weight_quantizer, self.weight_quantizer_0], [self.act_quantizer, self.act_quantizer_0]
        return [self.weight_quantizer], [self.act_quantizer]

    def _integer_path_ok(self, input):
        if not input.is_cuda or input.dtype != torch.float32 or torch.is_grad_enabled() and input.requires_grad:
            return False
        if input.numel() == 0:      # empty batch (a ragged last shard): no tile to launch; the elementwise + library route returns
            return False            # the empty result F.conv2d / F.linear give in the reference
        return self._integer_state_ok(input)

    def _integer_state_ok(self, input=None):
        """quantizer / layer state allows the integer path (input: only needed for the shape checks of split convs)"""
        if not (backend.integer_path and self.use_weight_quant and self.use_act_quant and not self.disable_act_quant):
            return False
        wqs, aqs = self._quantizers()
        for aq in aqs:
            if aq.inited is False or aq.delta is None or aq.channel_wise or (aq.is_training and aq.prob < 1.0):
                return False
            if torch.is_grad_enabled() and isinstance(aq.delta, nn.Parameter) and aq.delta.requires_grad and aq.is_training:
                return False
        for wq in wqs:
            if getattr(wq, 'inited', True) is False or wq.delta is None:
                return False
            if getattr(wq, 'soft_targets', False):
                return False
            if getattr(wq, 'round_mode', 'learned_hard_sigmoid') != 'learned_hard_sigmoid':
                return False
        if torch.is_grad_enabled() and any(getattr(wq, 'alpha', None) is not None and wq.alpha.requires_grad and
                                           getattr(wq, 'soft_targets', False) for wq in wqs):
            return False
        # hard limits of the kernels: u8 activation codes, s8 weight codes; anything wider takes the fake-quant route
        if any(aq.n_levels > 256 for aq in aqs) or any(wq.n_levels > 256 for wq in wqs):
            return False
        # 8-bit weight codes need the per-row activation code sums (zero-point fold), which exist per tensor, not per K range
        if self.split != 0 and any(wq.n_levels > 128 for wq in wqs):
            return False
        kw = self.fwd_kwargs
        if kw:
            if kw['groups'] != 1 or any(d != 1 for d in kw['dilation']):
                return False
            pad, stride = kw['padding'], kw['stride']
            if isinstance(pad, str) or len(set(pad)) != 1 or len(set(stride)) != 1:
                return False
            if self.fwd_func is F.conv1d and int(pad[0]) != 0:
                return False
            if self.split != 0 and input is not None and input.dim() == 4:
                # a split shortcut runs as two K-range GEMMs over the implicit (TMA-tiled) route only
                R, S = self.weight.shape[2], self.weight.shape[3]
                p0, s0 = int(pad[0]), int(stride[0])
                Ho = (input.shape[2] + 2 * p0 - R) // s0 + 1
                Wo = (input.shape[3] + 2 * p0 - S) // s0 + 1
                if s0 != 1 or not _implicit_tiling_ok(input.shape[0], Ho, Wo):
                    return False
        return True

    def _packed_weights(self):
        wqs, _ = self._quantizers()
        key = tuple((id(wq), wq.n_levels, _version_of(wq.delta), _version_of(wq.zero_point),
                     _version_of(getattr(wq, 'alpha', None))) for wq in wqs) + (_version_of(self.weight), self.split)
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        w = self.weight.detach()
        if w.dim() == 3:                       # conv1d [N, C, k]
            w = w.unsqueeze(2)
        elif w.dim() == 2:                     # linear [N, K]
            w = w.reshape(w.shape[0], w.shape[1], 1, 1)
        packs = []
        bounds = [(0, w.shape[1])] if self.split == 0 else [(0, self.split), (self.split, w.shape[1])]
        for wq, (c0, c1) in zip(wqs, bounds):
            alpha = getattr(wq, 'alpha', None)
            packs.append(ops.pack_weight(w, wq.delta, wq.zero_point, wq.n_levels, alpha=alpha, c_begin=c0, c_end=c1))
        self._packed = (key, packs)
        return packs

    def prenorm_fusable(self, x, norm) -> bool:
        """True when `self(norm(x))` can run as one normalise + quantize producer on the integer path: GroupNorm in front of a
        conv, LayerNorm in front of a linear; no hook may observe the normalised tensor or this module's input."""
        if isinstance(norm, nn.GroupNorm):
            shape_ok = (self.fwd_func is F.conv2d and x.dim() == 4) or (self.fwd_func is F.conv1d and x.dim() == 3)
        elif isinstance(norm, nn.LayerNorm):
            shape_ok = (self.fwd_func is F.linear and len(norm.normalized_shape) == 1 and
                        norm.normalized_shape[0] == x.shape[-1] and self.split == 0)
        else:
            return False
        return bool(backend.fuse_norm and shape_ok and self._integer_path_ok(x) and not self._forward_hooks
                    and not self._forward_pre_hooks and not norm._forward_hooks)

    def forward_prenorm(self, x, norm, silu: bool = True, scale=None, shift=None, split: int = 0, act_fn=F.silu,
                        residual=None, tokens_out: bool = False, bias_img=None, resample=None, emit=None):
        """`self(silu(norm(x) [* (1 + scale) + shift]))` for a GroupNorm in front of this conv (or a LayerNorm in front of this
        linear).  On the integer path normalisation + conditioning + SiLU + activation quantization run as ONE producer
        pass (edadm_gn_fold + edadm_norm_act_quant_nhwc, or edadm_layernorm_quant_rows); otherwise it is computed module
        by module like the reference does.  tokens_out (1x1 conv only): return [B, H*W, N] (the `b c h w -> b (h w) c`
        rearrangement of SpatialTransformer.forward) straight from the GEMM instead of NCHW."""
        if split != 0 and self.split == 0:
            self.split = split
            self.set_split()
        if isinstance(x, ops.CatPair) and not (self.prenorm_fusable(x, norm) and isinstance(norm, nn.GroupNorm) and resample is None
                                                and not tokens_out):
            x = x.materialize()
        if not self.prenorm_fusable(x, norm):
            if (backend.fuse_norm and isinstance(norm, nn.GroupNorm) and scale is None and split == 0 and self.split == 0
                    and residual is None and bias_img is None and resample is None and not tokens_out
                    and getattr(act_fn, '__name__', '') in ('silu', 'nonlinearity', '_swish') and self.fwd_func is F.conv2d
                    and not (self.use_act_quant and not self.disable_act_quant)
                    and not self._forward_hooks and not self._forward_pre_hooks and not norm._forward_hooks
                    and isinstance(self.activation_function, StraightThrough)):
                weight = self.weight_quantizer(self.weight) if self.use_weight_quant else self.org_weight
                bias = self.bias if self.use_weight_quant else self.org_bias
                if ops.conv3x3_small_n_ok(x, weight, self.fwd_kwargs):
                    # the UNet output head (GroupNorm -> SiLU -> conv to <= 4 channels, fp32 input by design of the reference):
                    # statistics by edadm_gn_fold, normalisation + SiLU applied while the stencil kernel loads its patches
                    a, s = ops.gn_fold(x, norm.weight, norm.bias, norm.num_groups, norm.eps)
                    self.last_path = 'fake' if self.use_weight_quant else 'fp'
                    return ops.conv3x3_small_n(x, weight, bias, affine=(a, s, _silu_mode(silu, act_fn)))
            h = norm(x)
            if scale is not None:
                h = h * (1 + scale) + shift
            if silu:
                h = act_fn(h)      # the block's own formulation of swish (x*sigmoid(x) in the DDIM UNet, nn.SiLU in LDM)
            if resample is not None:
                h = resample(h)
            out = self(h, split=split, residual=residual, bias_img=bias_img)
            return out.flatten(2).permute(0, 2, 1) if tokens_out else out
        self.last_path = 'int8'
        silu = _silu_mode(silu, act_fn)
        if isinstance(norm, nn.LayerNorm):
            assert scale is None and not silu
            if emit is not None:        # (codes, rowsum) of the consumer quantizer; the caller checked emit_ok()
                return self._forward_int8(x, rows=('layernorm', norm), emit=emit)
            return self._finish(self._forward_int8(x, rows=('layernorm', norm), residual=self._epilogue_residual(residual)), residual)
        a, s = ops.gn_fold(x, norm.weight, norm.bias, norm.num_groups, norm.eps, scale, shift)
        codes = None
        if resample is not None:
            # `resample` is the block's h_upd (openaimodel.py Upsample / Downsample without conv): 2x average pooling is fused
            # with the normalisation pass; nearest 2x upsampling commutes with the quantizer and is done on the u8 codes
            kind = _resample_kind(resample)
            if kind == 'down2' and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0:
                out = self.forward(ops.norm_act_pool2(x, a, s, silu), split=split, residual=residual, bias_img=bias_img)
                return out
            if kind == 'up2' and not self.split and not any(p.needs_rowsum for p in self._packed_weights()):
                _, aqs = self._quantizers()
                aq = ops.ActQuant(aqs[0].delta, aqs[0].zero_point, aqs[0].n_levels)
                pw0 = self._packed_weights()[0]
                q_lo, _ = ops.norm_act_quant_nhwc(x, a, s, silu, aq, 0, cp=pw0.Cp)
                pad = int(self.fwd_kwargs['padding'][0])
                codes = ops.upsample2x_codes(q_lo, x.shape[1], pad, aq)
            else:
                h = norm(x)
                if scale is not None:
                    h = h * (1 + scale) + shift
                if silu:
                    h = act_fn(h)
                return self(resample(h), split=split, residual=residual, bias_img=bias_img)
        fold = bias_img is not None and residual is None and not tokens_out and self._epilogue_residual(bias_img) is not None
        out = self._finish(self._forward_int8(x, affine=(a, s, silu), residual=self._epilogue_residual(residual),
                                              tokens_out=tokens_out, bias_img=bias_img if fold else None, codes=codes), residual)
        return out if bias_img is None or fold else out + bias_img.reshape(out.shape[0], out.shape[1], *([1] * (out.dim() - 2)))

    def forward_upsample2x(self, x):
        """`self(F.interpolate(x, scale_factor=2, mode="nearest"))` -- the conv of a resampling `Upsample` (openaimodel.py Upsample,
        ddim/models/diffusion.py Upsample).  Nearest upsampling commutes with the element-wise quantizer, so on the integer path the
        LOW-resolution tensor is quantized (a quarter of the elements) and the u8 codes are replicated straight into the
        halo-padded layout the GEMM reads; the 4x larger fp32 tensor never exists."""
        ok = (backend.fuse_norm and self.fwd_func is F.conv2d and x.dim() == 4 and not self.split and self._integer_path_ok(x)
              and not self._forward_hooks and not self._forward_pre_hooks and int(self.fwd_kwargs['stride'][0]) == 1
              and not any(p.needs_rowsum for p in self._packed_weights()))
        if not ok:
            return self(F.interpolate(x, scale_factor=2, mode="nearest"))
        self.last_path = 'int8'
        _, aqs = self._quantizers()
        aq = ops.ActQuant(aqs[0].delta, aqs[0].zero_point, aqs[0].n_levels)
        pw0 = self._packed_weights()[0]
        q_lo, _ = ops.act_quant_nhwc(x, aq, 0, cp=pw0.Cp)
        codes = ops.upsample2x_codes(q_lo, x.shape[1], int(self.fwd_kwargs['padding'][0]), aq)
        return self._finish(self._forward_int8(x, codes=codes), None)

    def forward_geglu(self, h, residual=None):
        """`self(a * gelu(g))` with (a, g) = h.chunk(2, -1): the GEGLU gate (ldm/modules/attention.py GEGLU.forward) folded
        into this linear's activation producer (edadm_geglu_quant_rows) on the integer path."""
        fusable = (backend.fuse_norm and self.fwd_func is F.linear and self.split == 0 and self._integer_path_ok(h)
                   and not self._forward_hooks and not self._forward_pre_hooks and h.shape[-1] == 2 * self.weight.shape[1])
        if not fusable:
            a, g = h.chunk(2, dim=-1)
            return self(a * F.gelu(g), residual=residual)
        self.last_path = 'int8'
        return self._finish(self._forward_int8(h, rows=('geglu',), residual=self._epilogue_residual(residual)), residual)

    def forward_from_tokens(self, y, hw, residual=None):
        """This 1x1 conv applied to tokens y [B, H*W, C] (the `b (h w) c -> b c h w` + proj_out of SpatialTransformer.forward):
        token rows ARE the NHWC layout the GEMM consumes, so no transpose pass is needed on the integer path."""
        H, W = hw
        B, T, C = y.shape
        ok = (self.fwd_func is F.conv2d and tuple(self.weight.shape[2:]) == (1, 1) and self.split == 0 and T == H * W
              and self._integer_path_ok(y) and not self._forward_hooks and not self._forward_pre_hooks
              and int(self.fwd_kwargs['padding'][0]) == 0 and int(self.fwd_kwargs['stride'][0]) == 1)
        if not ok:
            return self(y.permute(0, 2, 1).reshape(B, C, H, W), residual=residual)
        self.last_path = 'int8'
        return self._finish(self._forward_int8(y, rows=('tokens', H, W), residual=self._epilogue_residual(residual)), residual)

    def needs_act_rowsum(self) -> bool:
        """the integer GEMM of this layer needs per-row activation code sums (8-bit weight codes, zero-point fold)"""
        return any(p.needs_rowsum for p in self._packed_weights())

    def emit_ok(self, consumer_q, geglu: bool = False) -> bool:
        """This linear may emit the u8 codes of `consumer_q` (the next activation quantizer) from its GEMM epilogue: nothing but
        that quantizer (and, with `geglu`, the GEGLU gate) may observe its fp32 output.  The caller checks the input side
        (`prenorm_fusable` / `_integer_path_ok`)."""
        if not (backend.fuse_epilogue and self.fwd_func is F.linear and self.split == 0 and not self._forward_hooks
                and isinstance(self.activation_function, StraightThrough) and not torch.is_grad_enabled()):
            return False
        q = consumer_q
        if q.inited is False or q.delta is None or q.channel_wise or q.n_levels > 256 or (q.is_training and q.prob < 1.0):
            return False
        if any(p.w4 for p in self._packed_weights()):
            return False
        n_out = self.weight.shape[0] // 2 if geglu else self.weight.shape[0]
        return n_out % 32 == 0 if geglu else True

    def codes_consumer_ok(self) -> bool:
        """This linear can take its activation codes ready-made (from the producing GEMM's epilogue or a shared LayerNorm pass)."""
        return bool(self.fwd_func is F.linear and self.split == 0 and self._integer_state_ok() and not self._forward_hooks
                    and not self._forward_pre_hooks and not torch.is_grad_enabled())

    def forward_from_codes(self, q, rowsum, lead, residual=None, emit=None):
        """`self(x)` for an input that only exists as the u8 codes `q` [M, Kp] of this layer's own activation quantizer."""
        self.last_path = 'int8'
        if emit is not None:
            return self._forward_int8(None, rows=('codes', q, rowsum, lead), emit=emit)
        return self._finish(self._forward_int8(None, rows=('codes', q, rowsum, lead), residual=self._epilogue_residual(residual)), residual)

    def _post_ok(self, input, post):
        """the row-group term of edadm_qgemm_i8_rows_post applies: a plain integer-path linear, one weight pack, no row sums"""
        if self.fwd_func is not F.linear or self.split or not self._integer_path_ok(input) or input.dim() != 3:
            return False
        packs = self._packed_weights()
        N = packs[0].N
        return (len(packs) == 1 and not packs[0].needs_rowsum and not packs[0].w4 and N % 4 == 0 and post.dim() == 3
                and post.shape == (input.shape[0], 1, N) and post.dtype == torch.float32)

    def _epilogue_residual(self, residual):
        """`residual` if the GEMM epilogue may add it (nothing but a StraightThrough sits between conv and add)."""
        return residual if isinstance(self.activation_function, StraightThrough) else None

    def _finish(self, out, residual):
        out = self.activation_function(out)
        if residual is not None and self._epilogue_residual(residual) is None:
            out = out + residual
        return out

    def _forward_int8(self, input, affine=None, residual=None, rows=None, tokens_out=False, bias_img=None, codes=None, emit=None, post=None):
        """Exact integer GEMM: out = dA*dW[n]*sum (qa-za)(qw-zw) + bias  == the reference's fp32 conv of the
        dequantised tensors (quant_layer.py:414-434) without its per-product rounding."""
        packs = self._packed_weights()
        _, aqs = self._quantizers()
        aq = ops.ActQuant(aqs[0].delta, aqs[0].zero_point, aqs[0].n_levels,
                          self.split, aqs[1].delta if self.split else None,
                          aqs[1].zero_point if self.split else None, aqs[1].n_levels if self.split else 0)
        needs_rowsum = any(p.needs_rowsum for p in packs)
        if needs_rowsum and self.split:
            raise EdadmError("split shortcut with 8-bit weights is not supported on the integer path")
        bias = None if self.bias is None else self.bias.detach()
        pw0 = packs[0]
        N = pw0.N
        if rows is not None and rows[0] == 'tokens':
            # 1x1 conv over tokens [B, T, C]: rows of codes == NHWC codes; output NCHW (+ residual) straight from the GEMM
            _, H, W = rows
            B = input.shape[0]
            q, rowsum = ops.act_quant_rows(input.reshape(-1, input.shape[-1]), aq, want_rowsum=needs_rowsum)
            if backend.code_tap is not None:
                backend.code_tap(self, q, 0)
            out = torch.empty((B, N, H, W), dtype=torch.float32, device=input.device)
            if residual is not None:
                residual = residual.contiguous()
            self._gemm_chain(q, packs, aqs, out, H * W, bias, rowsum, residual)     # flat GEMM, NCHW store (out_hw = H*W)
            return out
        if self.fwd_func is F.linear:
            if rows is not None and rows[0] == 'codes':     # activation codes emitted by the producing GEMM / a shared LayerNorm pass
                _, q, rowsum, lead = rows
            else:
                lead = input.shape[:-1]
                if rows is None:
                    q, rowsum = ops.act_quant_rows(input.reshape(-1, input.shape[-1]), aq, want_rowsum=needs_rowsum)
                elif rows[0] == 'layernorm':
                    q, rowsum = ops.layernorm_quant_rows(input, rows[1].weight, rows[1].bias, rows[1].eps, aq, want_rowsum=needs_rowsum)
                else:   # 'geglu'
                    q, rowsum = ops.geglu_quant_rows(input, aq, want_rowsum=needs_rowsum)
            if backend.code_tap is not None:
                backend.code_tap(self, q, 0)
            if emit is not None:
                # the only consumer of this linear is the activation quantizer `emit[1]`: its codes come straight from the epilogue
                kind, cons, want_rs = emit
                return ops.qgemm_i8_codes(q, pw0, aqs[0].delta, aqs[0].zero_point, (cons.delta, cons.zero_point, cons.n_levels),
                                          bias=bias, rowsum=rowsum, geglu=(kind == 'geglu'), want_rowsum=want_rs)
            out = torch.empty((q.shape[0], N), dtype=torch.float32, device=q.device)
            if residual is not None:
                residual = residual.reshape(-1, N).contiguous()
            if q.shape[0] > 0 and post is not None:
                ops.qgemm_i8_rows_post(q, pw0, aqs[0].delta, aqs[0].zero_point, out, bias, residual, post.reshape(-1, N),
                                       q.shape[0] // post.shape[0])
            elif q.shape[0] > 0:
                self._gemm_chain(q, packs, aqs, out, 1, bias, rowsum, residual)
            return out.reshape(*lead, N)
        x4 = input.unsqueeze(2) if self.fwd_func is F.conv1d else input
        B, C, H, W = x4.shape
        R, S = pw0.R, pw0.S
        pad = int(self.fwd_kwargs['padding'][0])
        if codes is not None:      # activation codes prepared by the caller (halo included): [B, H + 2 pad, W + 2 pad, Cp]
            H, W = codes.shape[1] - 2 * pad, codes.shape[2] - 2 * pad
        stride = int(self.fwd_kwargs['stride'][0])
        pad_h = 0 if self.fwd_func is F.conv1d else pad
        if pad_h != pad:
            raise EdadmError("conv1d with padding is not supported on the integer path")
        Ho = (H + 2 * pad_h - R) // stride + 1
        Wo = (W + 2 * pad - S) // stride + 1
        cp_act = pw0.Cp if len(packs) == 1 else 0     # nibble-packed weights pad channels to 32: keep the im2col K aligned
        if codes is not None:
            q, chsum = codes, None
        elif affine is not None:
            q, chsum = ops.norm_act_quant_nhwc(x4, affine[0], affine[1], affine[2], aq, pad, want_chsum=needs_rowsum, cp=cp_act)
        else:
            q, chsum = ops.act_quant_nhwc(x4, aq, pad, want_chsum=needs_rowsum, cp=cp_act)
        rowsum = ops.conv_rowsum(chsum, Ho, Wo, R, S, stride) if needs_rowsum else None
        if backend.code_tap is not None:
            backend.code_tap(self, q, pad)
        if tokens_out and R == 1 and S == 1 and stride == 1 and pad == 0 and residual is None and not self.split:
            out = torch.empty((B, Ho * Wo, N), dtype=torch.float32, device=input.device)     # row-major [B*T][N] == tokens
            self._gemm_chain(q.reshape(-1, q.shape[-1]), packs, aqs, out, 1, bias, rowsum, None)
            return out
        out = torch.empty((B, N, Ho, Wo), dtype=torch.float32, device=input.device)
        if residual is not None:
            residual = residual.contiguous()
        if stride == 1 and _implicit_tiling_ok(B, Ho, Wo):
            self._gemm_chain(q, packs, aqs, out, Ho * Wo, bias, rowsum, residual, bias_img)
        else:
            if self.split:
                raise EdadmError("split shortcut on a strided / irregular conv is not supported on the integer path")
            a = ops.im2col_u8(q, Ho, Wo, R, S, stride)
            ops.qgemm_i8(a, pw0, aqs[0].delta, aqs[0].zero_point, out, Ho * Wo, bias=bias, rowsum=rowsum, filter_rs=(1, 1),
                         residual=residual, bias_img=bias_img)
        out = out.squeeze(2) if self.fwd_func is F.conv1d else out
        return out.flatten(2).permute(0, 2, 1) if tokens_out else out

    def _gemm_chain(self, q, packs, aqs, out, out_hw, bias, rowsum, residual=None, bias_img=None):
        if (len(packs) == 2 and rowsum is None and packs[0].Np == packs[1].Np and not any(p.w4 or p.needs_rowsum for p in packs)
                and backend.fuse_epilogue):
            # split shortcut: both K ranges in one launch, two TMEM accumulators combined in the epilogue
            ops.qgemm_i8_split(q, packs[0], packs[1], (aqs[0].delta, aqs[0].zero_point), (aqs[1].delta, aqs[1].zero_point), out, out_hw,
                               bias=bias, residual=residual, bias_img=bias_img)
            return
        c_off = 0
        last = len(packs) - 1
        for i, (pw, aqz) in enumerate(zip(packs, aqs)):
            ops.qgemm_i8(q, pw, aqz.delta, aqz.zero_point, out, out_hw, bias=bias if i == 0 else None, rowsum=rowsum,
                         a_c_offset=c_off, accumulate=i > 0, residual=residual if i == last else None,
                         bias_img=bias_img if i == last else None)
            c_off += pw.C

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, input: torch.Tensor, split: int = 0, residual=None, bias_img=None, post=None):
        """`residual` (optional, not in the reference signature): a tensor of the output's shape that is added to the
        result -- `conv(x) + residual` -- inside the GEMM epilogue on the integer path, as a plain add elsewhere.
        `bias_img` (optional, convs): [B, N(,1,1)] added per (image, channel) -- the ResBlock's `h + emb_out`.
        `post` (optional, linears with a residual): [B, 1, N], one row per sample added after the residual."""
        if isinstance(input, ops.CatPair) and not (self.fwd_func is F.conv2d and self._integer_path_ok(input)):
            input = input.materialize()
        if post is not None:
            if not (residual is not None and self._post_ok(input, post) and self._epilogue_residual(residual) is not None):
                return self.forward(input, split=split, residual=residual) + post
            self.last_path = 'int8'
            return self._finish(self._forward_int8(input, residual=residual, post=post), residual)
        if bias_img is not None:
            fold = (residual is None and self.fwd_func is F.conv2d and self._integer_path_ok(input)
                    and self._epilogue_residual(bias_img) is not None)
            if not fold:
                out = self.forward(input, split=split, residual=residual)
                return out + bias_img.reshape(out.shape[0], out.shape[1], *([1] * (out.dim() - 2)))
        if split != 0 and self.split != 0:
            assert split == self.split
        elif split != 0:
            logger.info(f"split at {split}!")
            self.split = split
            self.set_split()

        if self._integer_path_ok(input):
            self.last_path = 'int8'
            return self._finish(self._forward_int8(input, residual=self._epilogue_residual(residual), bias_img=bias_img), residual)

        if not self.disable_act_quant and self.use_act_quant:
            if self.split != 0:
                input = torch.cat([self.act_quantizer(input[:, :self.split, :, :]),
                                   self.act_quantizer_0(input[:, self.split:, :, :])], dim=1)
            else:
                input = self.act_quantizer(input)
        if self.use_weight_quant:
            if self.split != 0:
                weight = torch.cat([self.weight_quantizer(self.weight[:, :self.split, ...]),
                                    self.weight_quantizer_0(self.weight[:, self.split:, ...])], dim=1)
            else:
                weight = self.weight_quantizer(self.weight)
            bias = self.bias
        else:
            weight = self.org_weight
            bias = self.org_bias
        self.last_path = 'fake' if (self.use_weight_quant or self.use_act_quant) else 'fp'
        out = self.activation_function(_library_fwd(self.fwd_func, input, weight, bias, self.fwd_kwargs))
        return out if residual is None else out + residual


def _silu_mode(silu, act_fn) -> int:
    """activation code of the fused producers (csrc/pack.cu norm_act): 0 none, 1 `F.silu` / `nn.SiLU` (x / (1 + exp(-x))),
    2 `x * sigmoid(x)` (the DDIM UNet's nonlinearity), +16 SFU approximations"""
    if not silu:
        return 0
    mode = 2 if getattr(act_fn, '__name__', '') in ('nonlinearity', '_swish') else 1
    return mode + (16 if backend.fast_silu else 0)


def _swish(x):
    """x * sigmoid(x): the DDIM UNet's own spelling of SiLU (ddim/models/diffusion.py nonlinearity)"""
    return x * torch.sigmoid(x)


def _resample_kind(m):
    """'up2' / 'down2' for the conv-less 2x resampling modules of the LDM ResBlock (zoo or reference classes), else None."""
    op = getattr(m, 'op', None)
    if isinstance(op, nn.AvgPool2d):
        k, st = op.kernel_size, op.stride
        if (k in (2, (2, 2))) and (st in (2, (2, 2))) and op.padding in (0, (0, 0)) and not op.ceil_mode:
            return 'down2'
    if type(m).__name__ == 'Upsample' and getattr(m, 'use_conv', True) is False and getattr(m, 'dims', 2) == 2:
        return 'up2'
    return None


def _implicit_tiling_ok(B, Ho, Wo):
    """Mirror of the tile-shape checks in edadm_qgemm_i8: 128 consecutive output pixels form a W x H x B box."""
    if Wo >= 128:
        return Wo % 128 == 0
    if 128 % Wo:
        return False
    rows = 128 // Wo
    return (Ho % rows == 0) if rows <= Ho else (rows % Ho == 0)


def _library_fwd(fn, input, weight, bias, kwargs):
    """fp32 conv / linear of the calibration + FP paths (cuDNN / cuBLAS), TF32 off unless opted in.  The narrow output
    layer (<= 4 channels, its input is never quantized) takes the dedicated stencil kernel when no gradient is needed."""
    if input.numel() == 0:          # empty batch: the library's own empty result, no kernel of ours has a tile to run
        return fn(input, weight, bias, **kwargs)
    if fn is F.conv2d and ops.conv3x3_small_n_ok(input, weight, kwargs):
        return ops.conv3x3_small_n(input, weight, bias)
    if (fn is F.linear and backend.calib_gemm_bf16x3 and ops.linear_bf16x3_ok(input, weight) and
            (backend.in_recon or (torch.is_grad_enabled() and (input.requires_grad or weight.requires_grad)))):
        # the reconstruction loop's linears (forward, dgrad, wgrad): hand-written tcgen05 GEMM, bf16 x 3 split, fp32 accumulation
        return ops.linear_bf16x3(input, weight, bias)
    if fn is F.conv2d and backend.calib_gemm_bf16x3 and backend.in_recon and ops.conv_bf16x3_ok(input, weight, kwargs):
        # the reconstruction loop's stride-1 convolutions: forward, dgrad and wgrad on the same bf16 x 3 kernel
        ops.conv_wgrad_on_tensor_cores = backend.calib_conv_wgrad_bf16x3
        return ops.conv_bf16x3(input, weight, bias)
    if not input.is_cuda or backend.allow_tf32:
        return fn(input, weight, bias, **kwargs)
    prev_c, prev_m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return fn(input, weight, bias, **kwargs)
    finally:
        torch.backends.cudnn.allow_tf32 = prev_c
        torch.backends.cuda.matmul.allow_tf32 = prev_m
