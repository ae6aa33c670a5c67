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


def save_inp_oup_data(model: QuantModel, layer: Union[QuantModule, BaseQuantBlock], cali_data, asym: bool = False,
                      act_quant: bool = False, batch_size: int = 32, input_prob: bool = False, keep_gpu: bool = True,
                      batch_transform=None):
    """Returns (Resblock, cached_inps, cached_outs) laid out exactly like the reference (data_utils.py:67-75).
    `batch_transform` maps one calibration slice to the model's positional inputs (classifier-free-guidance variant)."""
    device = next(model.parameters()).device
    get_inp_out = GetLayerInpOut(model, layer, device=device, asym=asym, input_prob=input_prob, act_quant=act_quant)
    cali_data = qdist.shard_calibration(cali_data)
    store = (lambda t: t) if keep_gpu else (lambda t: t.cpu())
    inps, outs, syms, temb_inps, temb_syms = [], [], [], [], []
    resblock = False
    for i in range(int(cali_data[0].size(0) / batch_size)):
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
