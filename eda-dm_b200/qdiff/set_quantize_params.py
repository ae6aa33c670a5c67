"""Scale initialisation drivers (interface of the reference's qdiff/set_quantize_params.py:9-70): flip the
quantizers to un-inited, push calibration batches through the model so that each quantizer runs its
L2.4 range search on what it sees, then mark them inited."""
import logging
from typing import Union

import torch

from .quant_layer import QuantModule
from .quant_block import BaseQuantBlock, QuantAttnBlock, QuantQKMatMul, QuantSMVMatMul, QuantBasicTransformerBlock
from .quant_model import QuantModel

logger = logging.getLogger(__name__)


def _act_quantizers(root, ldm_matmuls: bool, transformer: bool = False):
    """Activation quantizers the reference (re-)initialises: QuantModule inputs (+ split twin), the q/k/v/w
    quantizers of QuantAttnBlock; on the LDM route also those of QuantQKMatMul / QuantSMVMatMul
    (set_quantize_params_LDM.py:31-36); on the conditional route also those of QuantBasicTransformerBlock
    (qdiff_control/set_quantize_params_Conditional.py:38-46)."""
    out = []
    for m in root.modules():
        if isinstance(m, QuantModule):
            out.append(m.act_quantizer)
            if m.split != 0:
                out.append(m.act_quantizer_0)
        if isinstance(m, QuantAttnBlock):
            out += [m.act_quantizer_k, m.act_quantizer_q, m.act_quantizer_v, m.act_quantizer_w]
        if ldm_matmuls and isinstance(m, QuantSMVMatMul):
            out += [m.act_quantizer_v, m.act_quantizer_w]
        if ldm_matmuls and isinstance(m, QuantQKMatMul):
            out += [m.act_quantizer_k, m.act_quantizer_q]
        if transformer and isinstance(m, QuantBasicTransformerBlock):
            for attn in (m.attn1, m.attn2):
                out += [attn.act_quantizer_q, attn.act_quantizer_k, attn.act_quantizer_v, attn.act_quantizer_w]
    return out


def _weight_quantizers(root, with_split_twin: bool):
    out = []
    for m in root.modules():
        if isinstance(m, QuantModule):
            out.append(m.weight_quantizer)
            if with_split_twin and m.split != 0:
                out.append(m.weight_quantizer_0)
    return out


def _to_device(t, device):
    return t.to(device, non_blocking=True) if torch.is_tensor(t) else t


def _model_device(module):
    return next(module.parameters()).device


def set_act_quantize_params(module: Union[QuantModel, QuantModule, BaseQuantBlock], cali_data, batch_size: int = 256,
                            all_attention: bool = False):
    """`all_attention=True` (not in the reference's signature) also covers the LDM matmul and transformer-block
    quantizers, for driving a bare UNet without the LatentDiffusion sampler wrappers."""
    logger.info("set_act_quantize_params")
    module.set_quant_state(True, True)
    for q in _act_quantizers(module, ldm_matmuls=all_attention, transformer=all_attention):
        q.set_inited(False)
    batch_size = min(batch_size, cali_data[0].size(0))
    device = _model_device(module)
    with torch.no_grad():
        for i in range(int(cali_data[0].size(0) / batch_size)):
            module(*[_to_device(_[i * batch_size:(i + 1) * batch_size], device) for _ in cali_data])
    for q in _act_quantizers(module, ldm_matmuls=all_attention, transformer=all_attention):
        q.set_inited(True)


def set_weight_quantize_params(model, cali_data):
    logger.info("set_weight_quantize_params")
    model.set_quant_state(True, False)
    for q in _weight_quantizers(model, with_split_twin=False):
        q.set_inited(False)
    batch_size = 32
    device = _model_device(model)
    with torch.no_grad():
        model(*[_to_device(_[:batch_size], device) for _ in cali_data])
    for q in _weight_quantizers(model, with_split_twin=True):
        q.set_inited(True)
