"""Reconstruction loop shared by block_reconstruction / layer_reconstruction (and their CFG twins).

Same optimisation problem, parameter set, optimiser, schedules and loss as the reference loop
(qdiff/block_recon.py:39-232, qdiff/layer_recon.py:39-129): AdaRound alphas of every QuantModule in the unit
(Adam, lr_w) and activation step sizes (Adam, lr_a), both cosine-annealed to 0; loss = L_p(block output,
FP output) + add_loss * sum of per-layer L2 terms between a quantized and an FP pass (FBR).

What is different on B200:
  * all quantizer arithmetic, its gradients and the loss reductions are the fused kernels of libedadm.so;
  * nothing in the loop synchronises with the host (losses stay on the device; the reference calls .cpu()
    on the loss every iteration, block_recon.py:186,194);
  * gradients live in one flat bucket (dist.GradBucket) and, when a process group is up, are averaged with a
    single all-reduce per step -- data parallel over the calibration rows of the unit;
  * the cache builder shards the calibration set by rank.
"""
import contextlib
import logging
import math
import random

import torch

from .quant_layer import QuantModule, lp_loss, backend
from .quant_block import (BaseQuantBlock, QuantAttnBlock, QuantAttentionBlock, QuantBasicTransformerBlock)
from .adaptive_rounding import AdaRoundQuantizer
from .utils import AttentionMap
from edadm import ops
from . import dist as qdist
from ._fused_adam import FusedAdam

logger = logging.getLogger(__name__)


class LinearTempDecay:
    """Temperature b of the rounding regulariser (reference block_recon.py:305-323)."""

    def __init__(self, t_max: int, rel_start_decay: float = 0.2, start_b: int = 10, end_b: int = 2):
        self.t_max = t_max
        self.start_decay = rel_start_decay * t_max
        self.start_b, self.end_b = start_b, end_b

    def __call__(self, t):
        if t < self.start_decay:
            return self.start_b
        rel_t = (t - self.start_decay) / (self.t_max - self.start_decay)
        return self.end_b + (self.start_b - self.end_b) * max(0.0, (1 - rel_t))


class LossFunction:
    """Reconstruction loss + optional rounding regulariser (reference block_recon.py:235-302).  Every caller in
    the reference passes round_loss='none'; 'relaxation' is kept and runs the fused regulariser kernel."""

    def __init__(self, block, round_loss: str = 'relaxation', weight: float = 1., rec_loss: str = 'mse',
                 max_count: int = 2000, b_range: tuple = (10, 2), decay_start: float = 0.0, warmup: float = 0.0,
                 p: float = 2.):
        self.block, self.round_loss, self.weight, self.rec_loss = block, round_loss, weight, rec_loss
        self.loss_start = max_count * warmup
        self.p, self.iters = p, max_count
        self.temp_decay = LinearTempDecay(max_count, rel_start_decay=warmup + (1 - warmup) * decay_start,
                                          start_b=b_range[0], end_b=b_range[1])
        self.count = 0

    def __call__(self, pred, tgt, grad=None):
        self.count += 1
        if self.rec_loss == 'mse':
            rec_loss = lp_loss(pred, tgt, p=self.p)
        elif self.rec_loss == 'fisher_diag':
            rec_loss = ((pred - tgt).pow(2) * grad.pow(2)).sum(1).mean()
        elif self.rec_loss == 'fisher_full':
            a, g = (pred - tgt).abs(), grad.abs()
            rec_loss = (torch.sum(a * g, (1, 2, 3)).view(-1, 1, 1, 1) * a * g).mean() / 100
        else:
            raise ValueError('Not supported reconstruction loss function: {}'.format(self.rec_loss))
        b = self.temp_decay(self.count)
        if self.count < self.loss_start or self.round_loss == 'none':
            return rec_loss
        if self.round_loss != 'relaxation':
            raise NotImplementedError
        round_loss = 0
        modules = [self.block] if isinstance(self.block, QuantModule) else self.block.modules()
        for module in modules:
            if isinstance(module, QuantModule):
                round_loss = round_loss + ops.round_reg(module.weight_quantizer.alpha, b, self.weight)
        return rec_loss + round_loss


def _as_param(q):
    q.delta = torch.nn.Parameter(q.delta.detach().clone())
    return q.delta


def _install_adaround(module: QuantModule, recon_w: bool, w_para: list, split_aware: bool = True):
    mode = 'learned_hard_sigmoid'
    if module.split == 0 or not split_aware:
        module.weight_quantizer = AdaRoundQuantizer(uaq=module.weight_quantizer, round_mode=mode,
                                                    weight_tensor=module.org_weight.data)
        quantizers = [module.weight_quantizer]
    else:
        module.weight_quantizer = AdaRoundQuantizer(uaq=module.weight_quantizer, round_mode=mode,
                                                    weight_tensor=module.org_weight.data[:, :module.split, ...].contiguous())
        module.weight_quantizer_0 = AdaRoundQuantizer(uaq=module.weight_quantizer_0, round_mode=mode,
                                                      weight_tensor=module.org_weight.data[:, module.split:, ...].contiguous())
        quantizers = [module.weight_quantizer, module.weight_quantizer_0]
    if recon_w:
        for q in quantizers:
            q.soft_targets = True
            w_para.append(q.alpha)


def _attention_quantizers(module, transformer: bool):
    """q/k/v/w quantizers that become trainable, in the reference's order."""
    if isinstance(module, QuantAttentionBlock) and not transformer:
        a = module.attention
        return [a.qkv_matmul.act_quantizer_q, a.qkv_matmul.act_quantizer_k, a.smv_matmul.act_quantizer_v,
                a.smv_matmul.act_quantizer_w]
    if isinstance(module, QuantAttnBlock):
        return [module.act_quantizer_q, module.act_quantizer_k, module.act_quantizer_v, module.act_quantizer_w]
    if isinstance(module, QuantBasicTransformerBlock) and transformer:
        out = []
        for attn in (module.attn1, module.attn2):
            out += [attn.act_quantizer_q, attn.act_quantizer_k, attn.act_quantizer_v, attn.act_quantizer_w]
        return out
    return []


def prepare_unit(unit, act_quant, recon_w, recon_a, transformer=False, with_hooks=True, split_aware=True,
                 attn_only=False):
    """Swap in AdaRound quantizers, promote step sizes to Parameters; returns (w_para, a_para, hooks, trained).
    attn_only: only the q/k/v/w quantizers of a QuantAttnBlock are trained (reference attn_layer_recon.py:42-58)."""
    w_para, a_para, hooks, trained = [], [], [], []
    modules = [unit] if isinstance(unit, QuantModule) else list(unit.modules())
    for module in modules:
        if isinstance(module, QuantModule) and not attn_only:
            if with_hooks:
                hooks.append(AttentionMap(module))
            _install_adaround(module, recon_w, w_para, split_aware)
        if isinstance(module, (QuantModule, BaseQuantBlock)) and act_quant:
            quantizers = list(_attention_quantizers(module, transformer))
            if attn_only:
                quantizers = quantizers if isinstance(module, QuantAttnBlock) else []
            elif module.act_quantizer.delta is not None:
                quantizers.append(module.act_quantizer)
                if module.split != 0 and isinstance(module, QuantModule):
                    quantizers.append(module.act_quantizer_0)
            for q in quantizers:
                p = _as_param(q)
                if recon_a:
                    a_para.append(p)
                    q.is_training = True
                    trained.append(q)
    return w_para, a_para, hooks, trained


def finish_unit(unit, trained, hooks, attn_only=False):
    modules = [unit] if isinstance(unit, QuantModule) else list(unit.modules())
    for module in modules:
        if isinstance(module, QuantModule) and not attn_only:
            module.weight_quantizer.soft_targets = False
            module.act_quantizer.is_training = False
            if module.split != 0:
                if hasattr(module.weight_quantizer_0, 'soft_targets'):
                    module.weight_quantizer_0.soft_targets = False
                module.act_quantizer_0.is_training = False
    for q in trained:
        q.is_training = False
    for hook in hooks:
        hook.remove()


def _take(t, idx, device):
    out = t[idx]
    return out if out.device == device else out.to(device, non_blocking=True)


def _cosine_lr(lr0, t, t_max):
    """Closed form of CosineAnnealingLR(T_max=t_max, eta_min=0) after t scheduler steps."""
    return lr0 * (1.0 + math.cos(math.pi * min(t, t_max) / t_max)) / 2.0


def _memoise_fp_taps(unit, hooks, sources, resblock, sz, batch, act_quant, max_bytes):
    """FP forward of every cached sample (inputs: the FP-model activations sources[2] / sources[4]); returns one [sz, ...]
    tensor per hooked QuantModule but the last (the reference leaves it out of the per-layer loss, block_recon.py:188), or None
    when they would not fit in `max_bytes`."""
    out = None
    unit.set_quant_state(False, False)
    prev, backend.in_recon = backend.in_recon, True          # the same kernels the loop itself would run
    try:
        with torch.no_grad():
            for i in range(0, sz, batch):
                j = min(sz, i + batch)
                args = (sources[2][i:j], sources[4][i:j]) if resblock else (sources[2][i:j],)
                unit(*args)
                taps = [h.out for h in hooks][:-1]
                if out is None:
                    if not taps or not all(torch.is_tensor(t) and t.shape[0] == j - i for t in taps):
                        return None
                    if sum(t[0].numel() * t.element_size() for t in taps) * sz > max_bytes:
                        return None
                    out = [torch.empty((sz,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in taps]
                for dst, t in zip(out, taps):
                    dst[i:j].copy_(t)
    finally:
        backend.in_recon = prev
        unit.set_quant_state(True, act_quant)
    return out


def _graph_capturable(unit, device):
    if device.type != 'cuda' or not backend.recon_cuda_graph:
        return False
    if backend.qdrop_inkernel_rng:
        # the Philox (seed, offset) of in-kernel QDrop draws are host scalars: a captured launch would replay ONE mask forever
        return False
    # Checkpointed units (QuantBasicTransformerBlock, QuantAttentionBlock: qdiff.quant_block._Recompute) are captured too: the
    # recompute is a nested autograd pass on the capturing stream, and its QDrop re-draws come from torch's graph-safe Philox
    # generator, so every replay sees fresh masks exactly like the eager loop (backend.recon_graph_checkpointed = False opts out).
    if not backend.recon_graph_checkpointed:
        for m in unit.modules():
            if isinstance(m, QuantAttentionBlock) or getattr(m, 'use_checkpoint', False) or getattr(m, 'checkpoint', False) is True:
                return False
    return True


def reconstruct(model, unit, cali_data, *, batch_size, iters, weight, opt_mode, asym, b_range, warmup, act_quant, lr_a,
                lr_w, p, input_prob, keep_gpu, recon_w, recon_a, add_loss, cache_builder, cache_batch_size=32,
                transformer=False, is_layer=False, split_aware=True, attn_only=False, return_losses=False, timing=None):
    """One reconstruction unit.  After three eager iterations the whole step -- input mixing, the (up to three)
    forwards, both losses, backward, the gradient all-reduce and both Adam updates -- is captured in ONE CUDA graph and
    replayed; per iteration the host only draws the minibatch indices, gathers the cached rows into static buffers and
    writes the two cosine learning rates."""
    unit.set_quant_state(True, act_quant)
    w_para, a_para, hooks, trained = prepare_unit(unit, act_quant, recon_w, recon_a, transformer,
                                                  with_hooks=not (is_layer or attn_only), split_aware=split_aware,
                                                  attn_only=attn_only)
    device = next(model.parameters()).device
    use_graph = _graph_capturable(unit, device) and bool(w_para or a_para)
    w_opt = a_opt = w_sched = a_sched = None
    w_lr = a_lr = None
    fused_adam = use_graph and backend.recon_fused_adam
    if use_graph:   # learning rates live on the device so the captured Adam steps see the schedule
        lrs = torch.tensor([float(lr_w), float(lr_a)], device=device)
        w_lr, a_lr = lrs[0], lrs[1]
        if w_para and not fused_adam:
            w_opt = torch.optim.Adam(w_para, lr=w_lr, capturable=True)
        if a_para and not fused_adam:
            a_opt = torch.optim.Adam(a_para, lr=a_lr, capturable=True)
    else:
        if w_para:
            w_opt = torch.optim.Adam(w_para, lr=lr_w)
            w_sched = torch.optim.lr_scheduler.CosineAnnealingLR(w_opt, T_max=iters, eta_min=0.)
        if a_para:
            a_opt = torch.optim.Adam(a_para, lr=lr_a)
            a_sched = torch.optim.lr_scheduler.CosineAnnealingLR(a_opt, T_max=iters, eta_min=0.)
    loss_func = LossFunction(unit, round_loss='none', weight=weight, max_count=iters, rec_loss=opt_mode,
                             b_range=b_range, decay_start=0, warmup=warmup, p=p)

    resblock, cached_inps, cached_outs = cache_builder(model, unit, cali_data, asym, act_quant,
                                                       batch_size=cache_batch_size, input_prob=True, keep_gpu=keep_gpu)
    sz = cached_outs.size(0)
    model.block_count = model.block_count + 1
    bucket = qdist.GradBucket(w_para + a_para) if (w_para or a_para) else None
    # both Adam updates as ONE pass over the flat gradient bucket (edadm_fused_adam), which also clears the consumed gradients
    adam = FusedAdam(bucket, len(w_para), lrs) if (fused_adam and bucket is not None) else None
    rng = random if not qdist.is_active() else random.Random(random.getrandbits(48) + 7919 * qdist.rank())
    losses = []
    fbr = (not is_layer) and len(hooks) != 0 and add_loss != 0.0
    bsz = min(batch_size, sz)

    # cache tensors in the order (out, inp, sym[, emb_inp, emb_sym])
    if resblock:
        sources = [cached_outs, cached_inps[0][0], cached_inps[1][0], cached_inps[0][1], cached_inps[1][1]]
    else:
        sources = [cached_outs, cached_inps[0], cached_inps[1]]
    on_device = all(t.device == device for t in sources)
    # The FP forward of the per-layer terms (reference :161-165) is a deterministic function of the cached sample: with the cache
    # in HBM its taps are computed ONCE per unit for all samples and gathered per iteration like the cached block output,
    # instead of being recomputed in each of the 20 000 iterations (backend.recon_memoise_fp_taps, bounded by
    # backend.recon_memoise_bytes).
    n_base = len(sources)
    memo = False
    if fbr and on_device and backend.recon_memoise_fp_taps and device.type == 'cuda':
        taps = _memoise_fp_taps(unit, hooks, sources, resblock, sz, cache_batch_size, act_quant, backend.recon_memoise_bytes)
        if taps is not None:
            sources = sources + taps
            memo = True
    # Device-resident iteration state: with the cache in HBM the whole iteration -- minibatch gather included -- reads its
    # indices and learning rates from tables indexed by a device-side counter, so a captured step needs NO per-iteration host
    # work (the host only replays; 8 data-parallel ranks stay in lockstep instead of waiting for the slowest host loop at the
    # all-reduce).  The index table is drawn up front with the reference's own call sequence (one random.sample per iteration).
    device_state = use_graph and on_device
    it_counter = torch.zeros(1, dtype=torch.long, device=device)
    if device_state:
        idx_all = torch.tensor([rng.sample(range(sz), bsz) for _ in range(iters)], dtype=torch.long, device=device)
        w_lr_all = torch.tensor([_cosine_lr(lr_w, t, iters) for t in range(iters)], dtype=torch.float32, device=device)
        a_lr_all = torch.tensor([_cosine_lr(lr_a, t, iters) for t in range(iters)], dtype=torch.float32, device=device)
        loss_all = torch.zeros(iters, dtype=torch.float32, device=device)
    static = [torch.empty((bsz,) + tuple(t.shape[1:]), dtype=t.dtype, device=device) for t in sources] \
        if (use_graph and not device_state) else None
    loss_out = torch.zeros((), dtype=torch.float32, device=device)
    gnorm_out = torch.zeros(2, dtype=torch.float32, device=device)      # diagnostics: |grad| of the alpha / step-size parameters
    n_w = sum(p_.numel() for p_ in w_para)

    def gather(idx):
        idx_t = torch.as_tensor(idx, device=sources[0].device)
        rows = [_take(t, idx_t, device) for t in sources]
        if static is None:
            return rows
        for dst, src in zip(static, rows):
            dst.copy_(src, non_blocking=True)
        return static

    def device_rows():
        """this iteration's minibatch and learning rates, selected on the device by the iteration counter"""
        idx = idx_all.index_select(0, it_counter).reshape(-1)
        w_lr.copy_(w_lr_all.index_select(0, it_counter).reshape(()))
        a_lr.copy_(a_lr_all.index_select(0, it_counter).reshape(()))
        return [t.index_select(0, idx) for t in sources]

    def step(rows):
        cur_out, cur_inp, cur_sym = rows[0], rows[1], rows[2]
        emb_inp, emb_sym = (rows[3], rows[4]) if resblock else (None, None)
        if input_prob < 1.0:
            cur_inp = torch.where(torch.rand_like(cur_inp) < input_prob, cur_inp, cur_sym)
        elif not is_layer:
            cur_inp = cur_sym
        if bucket is not None:
            bucket.zero(memset=adam is None)
        args_q = (cur_inp, emb_inp) if resblock else (cur_inp,)
        args_fp = (cur_sym, emb_sym) if resblock else (cur_sym,)
        m_loss = 0.0
        if fbr and memo:
            out_quant = unit(*args_q)
            unit(*args_q)
            module_q = [h.out for h in hooks]
            module_r = list(rows[n_base:]) + [None]
        elif fbr and fp_stream is not None:
            # The FP forward (targets of the per-layer terms, no gradient) does not depend on the two quantized forwards: it
            # is issued on a forked stream so that, inside the captured graph, its many small kernels overlap with theirs.
            # Python order == reference order (quantized, FP, quantized); only the stream assignment differs.
            main = torch.cuda.current_stream(device)
            fp_stream.wait_stream(main)
            out_quant = unit(*args_q)
            unit.set_quant_state(False, False)
            with torch.cuda.stream(fp_stream), torch.no_grad():
                unit(*args_fp)
            module_r = [h.out for h in hooks]
            for t in module_r:
                if torch.is_tensor(t):
                    t.record_stream(main)
            unit.set_quant_state(True, act_quant)
            unit(*args_q)
            module_q = [h.out for h in hooks]
            main.wait_stream(fp_stream)
        else:
            out_quant = unit(*args_q)
        if fbr and fp_stream is None and not memo:
            unit.set_quant_state(False, False)
            with torch.no_grad():
                unit(*args_fp)
            module_r = [h.out for h in hooks]
            unit.set_quant_state(True, act_quant)
            unit(*args_q)
            module_q = [h.out for h in hooks]
        if fbr:
            for j in range(len(module_r) - 1):      # the unit's last QuantModule is left out (reference :188)
                m_loss = m_loss + lp_loss(module_q[j], module_r[j], p=2)
        block_loss = loss_func(out_quant, cur_out)
        loss = block_loss + add_loss * m_loss if fbr else block_loss
        loss.backward()
        if bucket is not None:
            bucket.all_reduce_mean()
            if timing is not None and 'grad_norms' in timing:
                gnorm_out[0].copy_(bucket.flat[:n_w].norm()); gnorm_out[1].copy_(bucket.flat[n_w:].norm())
        if adam is not None:
            adam.step()
        for opt in (w_opt, a_opt):
            if opt is not None:
                opt.step()
        loss_out.copy_(block_loss.detach())
        if device_state:
            loss_all.index_copy_(0, it_counter, block_loss.detach().reshape(1))
            it_counter.add_(1)

    graph = None
    n_eager = 3
    # Warm-up iterations and the capture run on ONE non-default stream (library handles / workspaces used by the
    # autograd thread must never have been bound to the legacy stream, or capture is invalidated).
    side = torch.cuda.Stream(device=device) if use_graph else None
    if bucket is not None and device.type == 'cuda' and backend.recon_overlap_allreduce:
        bucket.overlap_backward(stream=side)
    # a parallel FP branch pays off while the unit's kernels are launch/latency bound (+4 % on a church 16x16 ResBlock) and
    # hurts on larger units (ImageNet 32x32 ResBlock: -23 %): opt-in (backend.recon_overlap_fp) and size-gated
    small_unit = use_graph and bsz * sources[1][0].numel() <= (4 << 20)
    fp_stream = torch.cuda.Stream(device=device) if (use_graph and backend.recon_overlap_fp and small_unit) else None
    if side is not None:
        side.wait_stream(torch.cuda.current_stream(device))
    stream_ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
    backend.in_recon = True
    with stream_ctx:
        for it in range(iters):
            if use_graph and graph is None and it == n_eager:
                try:
                    side.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=side):
                        step(device_rows() if device_state else static)
                except Exception as exc:  # keep optimising eagerly; the captured work never ran
                    logger.warning("CUDA-graph capture of the reconstruction step failed (%s); continuing eagerly", exc)
                    graph, use_graph = False, False
            if timing is not None and it == timing.get('warmup', 0):
                timing['start'], timing['end'] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                qdist.barrier()
                torch.cuda.synchronize()
                timing['start'].record()
            prof = timing is not None and timing.get('profile_iter') == it
            if prof:                                    # ncu --profile-from-start off: exactly this iteration is captured
                torch.cuda.synchronize(); torch.cuda.profiler.start()
            if device_state:
                if graph:
                    graph.replay()
                else:
                    step(device_rows())
            else:
                rows = gather(rng.sample(range(sz), bsz))
                if w_lr is not None:
                    w_lr.fill_(_cosine_lr(lr_w, it, iters))
                    a_lr.fill_(_cosine_lr(lr_a, it, iters))
                if graph:
                    graph.replay()
                else:
                    step(rows)
                    for sched in (w_sched, a_sched):
                        if sched is not None:
                            sched.step()
                if return_losses:
                    losses.append(loss_out.clone())
            if prof:
                torch.cuda.synchronize(); torch.cuda.profiler.stop()
            if timing is not None and 'grad_norms' in timing:
                timing['grad_norms'].append(gnorm_out.clone())
        if timing is not None and 'start' in timing:
            timing['end'].record()
    backend.in_recon = False
    if side is not None:
        torch.cuda.current_stream(device).wait_stream(side)

    if timing is not None and 'start' in timing:
        torch.cuda.synchronize()
        qdist.barrier()
        timing['iters'] = iters - timing.get('warmup', 0)
        timing['ms_per_iter'] = timing['start'].elapsed_time(timing['end']) / max(1, timing['iters'])
        timing['bucket_bytes'] = bucket.nbytes() if bucket is not None else 0
        timing['cuda_graph'] = bool(graph)
        timing['fp_taps_memoised'] = memo
    finish_unit(unit, trained, hooks, attn_only)
    if adam is not None:
        adam.finish()
    if bucket is not None:
        bucket.release()
        for prm in bucket.params:
            prm.grad = None
    if return_losses:
        if device_state:
            return loss_all.clone()
        return torch.stack(losses) if losses else torch.empty(0)
    return None
