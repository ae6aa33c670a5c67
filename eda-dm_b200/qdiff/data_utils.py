"""Calibration cache builder (interface of the reference's qdiff/data_utils.py:7-171): for one block / layer,
run the network prefix on every calibration batch and keep (quantized-path input, FP output, FP-path input).

B200 notes: the prefix forwards of the quantized network run on the integer tcgen05 path (no_grad, hard
rounding); under torch.distributed each rank builds only its shard of the cache (rows rank::world of the
calibration set), so cache construction scales with the number of GPUs and the cache itself stays in HBM.
"""
from typing import Union

import torch

from .quant_layer import QuantModule
from .quant_model import QuantModel
from .quant_block import BaseQuantBlock
from . import dist as qdist


class StopForwardException(Exception):
    """Raised by the hook to stop the forward pass once the unit of interest has run."""


class DataSaverHook:
    def __init__(self, store_input=False, store_output=False, stop_forward=False):
        self.store_input, self.store_output, self.stop_forward = store_input, store_output, stop_forward
        self.input_store = None
        self.output_store = None

    def __call__(self, module, input_batch, output_batch):
        if self.store_input:
            self.input_store = input_batch
        if self.store_output:
            self.output_store = output_batch
        if self.stop_forward:
            raise StopForwardException


def _detach_inputs(store):
    if len(store) == 1:
        return store[0].detach()
    return (store[0].detach(), store[1].detach())


class GetLayerInpOut:
    def __init__(self, model: QuantModel, layer: Union[QuantModule, BaseQuantBlock], device: torch.device,
                 input_prob: bool = False, act_quant: bool = False, asym: bool = False):
        self.model, self.layer, self.device = model, layer, device
        self.asym, self.act_quant, self.input_prob = asym, act_quant, input_prob
        self.data_saver = DataSaverHook(store_input=True, store_output=True, stop_forward=True)

    def _run(self, model_input):
        try:
            self.model(*[_.to(self.device) for _ in model_input])
        except StopForwardException:
            pass

    def __call__(self, model_input):
        self.model.eval()
        self.model.set_quant_state(False, False)
        handle = self.layer.register_forward_hook(self.data_saver)
        input_sym = None
        with torch.no_grad():
            self._run(model_input)                      # FP network: FP input and FP output of the unit
            if self.input_prob:
                input_sym = _detach_inputs(self.data_saver.input_store)
            if self.asym:                               # quantized network: what the unit will really be fed
                self.data_saver.store_output = False
                self.model.set_quant_state(weight_quant=True, act_quant=self.act_quant)
                self._run(model_input)
            self.data_saver.store_output = True
        handle.remove()
        input_store = _detach_inputs(self.data_saver.input_store)
        resblock = isinstance(input_store, tuple)
        if self.input_prob:
            return resblock, input_store, self.data_saver.output_store.detach(), input_sym
        return resblock, input_store, self.data_saver.output_store.detach()


class StagedCache:
    """Prefix reuse for the cache builder.

    The reference re-runs the network from its input up to the unit -- once with FP weights, once fully quantized -- for every
    calibration batch of EVERY unit (data_utils.py:125-171), i.e. O(units^2) prefix work over a whole-model reconstruction.  The
    UNet structures in unet_zoo expose their forward pass as a list of stages over an explicit state (h, the skip stack, the
    embedding); this cache keeps, for every calibration batch, the FP-path and the quantized-path state at the frontier stage of
    the units handled so far, resident in HBM (fp32, <= 17 MB per ImageNet sample and path).  A unit is then served by running
    only ITS stage from the stored state; every stage of every path runs once per batch over the whole walk.  The tensors are
    the ones the reference would compute: same ops, same order, same quantizer state (units behind the frontier are final).
    The frontier states are dropped and rebuilt whenever a quantizer or weight behind the frontier changed after it was passed."""

    def __init__(self, model, cali_data, batch_size, act_quant, batch_transform):
        self.model, self.unet = model, model.model
        self.batch_size, self.act_quant, self.batch_transform = batch_size, act_quant, batch_transform
        self.key = self.make_key(cali_data, batch_size, act_quant, batch_transform)
        self.cali = cali_data
        self.stage_mods = self.unet.stage_modules()
        self.stage_of = {}
        for k, mods in enumerate(self.stage_mods):
            for m in mods:
                for sub in m.modules():
                    self.stage_of[id(sub)] = k
        self.reset()

    @staticmethod
    def make_key(cali_data, batch_size, act_quant, batch_transform):
        return (tuple((t.data_ptr(), tuple(t.shape)) for t in cali_data), batch_size, bool(act_quant), batch_transform)

    def reset(self):
        self.stage = 0
        self.fp_states = self.q_states = None
        self.signature = None

    def _prefix_signature(self, upto):
        sig = []
        for mods in self.stage_mods[:upto]:
            for m in mods:
                for sub in m.modules():
                    for name in ('delta', 'alpha', 'zero_point', 'weight'):
                        t = getattr(sub, name, None)
                        if torch.is_tensor(t):
                            sig.append((id(sub), name, t.data_ptr(), t._version))
                    if hasattr(sub, 'soft_targets'):
                        sig.append((id(sub), 'soft', bool(sub.soft_targets)))
                    if hasattr(sub, 'inited') and hasattr(sub, 'n_bits'):
                        sig.append((id(sub), 'q', bool(sub.inited), sub.n_bits))
        return tuple(sig)

    def _batches(self):
        n = int(self.cali[0].size(0) / self.batch_size)
        device = next(self.model.parameters()).device
        for i in range(n):
            batch = [_[i * self.batch_size:(i + 1) * self.batch_size] for _ in self.cali]
            if self.batch_transform is not None:
                batch = self.batch_transform(batch)
            yield [_.to(device) for _ in batch]

    def _advance(self, target, asym):
        if self.fp_states is None:
            self.fp_states = [self.unet.stage_begin(*b) for b in self._batches()]
            self.q_states = [dict(st) for st in self.fp_states] if asym else None
            self.stage = 0
        elif self.signature != self._prefix_signature(self.stage):
            self.reset()
            return self._advance(target, asym)
        with torch.no_grad():
            while self.stage < target:
                k = self.stage
                self.model.set_quant_state(False, False)
                self.fp_states = [self.unet.run_stage(k, st) for st in self.fp_states]
                if self.q_states is not None:
                    self.model.set_quant_state(weight_quant=True, act_quant=self.act_quant)
                    self.q_states = [self.unet.run_stage(k, st) for st in self.q_states]
                self.stage += 1
        self.signature = self._prefix_signature(self.stage)

    def _run_unit_stage(self, k, st, saver, layer):
        handle = layer.register_forward_hook(saver)
        try:
            self.unet.run_stage(k, st)
        except StopForwardException:
            pass
        finally:
            handle.remove()

    def collect(self, layer, asym, input_prob):
        """per calibration batch: (resblock, input [quantized path when asym], FP output[, FP-path input]) like GetLayerInpOut"""
        k = self.stage_of[id(layer)]
        if self.q_states is None and asym and self.fp_states is not None:
            self.reset()
        if self.stage > k:
            self.reset()
        self._advance(k, asym)
        self.model.eval()
        out = []
        with torch.no_grad():
            for b in range(len(self.fp_states)):
                saver = DataSaverHook(store_input=True, store_output=True, stop_forward=True)
                self.model.set_quant_state(False, False)
                self._run_unit_stage(k, self.fp_states[b], saver, layer)
                fp_out = saver.output_store.detach()
                input_sym = _detach_inputs(saver.input_store)
                inp = input_sym
                if asym:
                    saver = DataSaverHook(store_input=True, store_output=False, stop_forward=True)
                    self.model.set_quant_state(weight_quant=True, act_quant=self.act_quant)
                    self._run_unit_stage(k, self.q_states[b], saver, layer)
                    if saver.input_store is None:
                        raise RuntimeError(f"{type(layer).__name__} did not run in stage {k} of the quantized network")
                    inp = _detach_inputs(saver.input_store)
                resblock = isinstance(inp, tuple)
                out.append((resblock, inp, fp_out, input_sym) if input_prob else (resblock, inp, fp_out))
        return out


def _staged_cache_for(model, layer, cali_data, batch_size, act_quant, batch_transform):
    """the model's StagedCache if prefix reuse applies to this unit, else None"""
    from .quant_layer import backend
    unet = getattr(model, 'model', None)
    if not backend.cache_prefix_reuse or unet is None or not hasattr(unet, 'run_stage') or not hasattr(unet, 'stage_modules'):
        return None
    cache = getattr(model, '_stage_cache', None)
    key = StagedCache.make_key(cali_data, batch_size, act_quant, batch_transform)
    if cache is None or cache.key != key:
        cache = StagedCache(model, cali_data, batch_size, act_quant, batch_transform)
        object.__setattr__(model, '_stage_cache', cache)
    return cache if id(layer) in cache.stage_of else None


def save_inp_oup_data(model: QuantModel, layer: Union[QuantModule, BaseQuantBlock], cali_data, asym: bool = False,
                      act_quant: bool = False, batch_size: int = 32, input_prob: bool = False, keep_gpu: bool = True,
                      batch_transform=None):
    """Returns (Resblock, cached_inps, cached_outs) laid out exactly like the reference (data_utils.py:67-75).
    `batch_transform` maps one calibration slice to the model's positional inputs (classifier-free-guidance variant)."""
    device = next(model.parameters()).device
    get_inp_out = GetLayerInpOut(model, layer, device=device, asym=asym, input_prob=input_prob, act_quant=act_quant)
    cali_data = qdist.shard_calibration(cali_data)
    staged = _staged_cache_for(model, layer, cali_data, batch_size, act_quant, batch_transform)
    staged_results = staged.collect(layer, asym, input_prob) if staged is not None else None
    store = (lambda t: t) if keep_gpu else (lambda t: t.cpu())
    inps, outs, syms, temb_inps, temb_syms = [], [], [], [], []
    resblock = False
    for i in range(int(cali_data[0].size(0) / batch_size)):
        if staged_results is not None:
            res = staged_results[i]
        else:
            batch = [_[i * batch_size:(i + 1) * batch_size] for _ in cali_data]
            res = get_inp_out(batch_transform(batch) if batch_transform is not None else batch)
        resblock, cur_inp, cur_out = res[0], res[1], res[2]
        cur_sym = res[3] if input_prob else None
        if resblock:
            inps.append(store(cur_inp[0])); temb_inps.append(store(cur_inp[1]))
            if input_prob:
                syms.append(store(cur_sym[0])); temb_syms.append(store(cur_sym[1]))
        else:
            inps.append(store(cur_inp))
            if input_prob:
                syms.append(store(cur_sym))
        outs.append(store(cur_out))

    def cat(parts):
        t = torch.cat(parts)
        return t if keep_gpu or not torch.cuda.is_available() else t.pin_memory()

    cached_inps, cached_outs = cat(inps), cat(outs)
    if input_prob:
        cached_sym = cat(syms)
        if resblock:
            return resblock, ([cached_inps, cat(temb_inps)], [cached_sym, cat(temb_syms)]), cached_outs
        return resblock, (cached_inps, cached_sym), cached_outs
    if resblock:
        return resblock, ([cached_inps, cat(temb_inps)]), cached_outs
    return resblock, (cached_inps,), cached_outs
