"""QuantModel: in-place rewrite of a diffusion UNet into its quantized form, with the reference's
constructor and methods (qdiff/quant_model.py:12-95)."""
import logging

import torch
import torch.nn as nn

from .quant_block import get_specials, BaseQuantBlock
from .quant_block import QuantBasicTransformerBlock, QuantResBlock  # noqa: F401
from .quant_block import QuantQKMatMul, QuantSMVMatMul, QuantAttnBlock
from .quant_layer import QuantModule, UniformAffineQuantizer, StraightThrough
from . import attention as qattn

logger = logging.getLogger(__name__)


class QuantModel(nn.Module):

    def __init__(self, model: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}, **kwargs):
        super().__init__()
        self.model = model
        self.block_count = 0
        self.sm_abit = kwargs.get('sm_abit', 8)
        self.in_channels = model.in_channels
        if hasattr(model, 'image_size'):
            self.image_size = model.image_size
        self.specials = get_specials(act_quant_params['leaf_param'])
        self.quant_module_refactor(self.model, weight_quant_params, act_quant_params)
        self.quant_block_refactor(self.model, weight_quant_params, act_quant_params)
        self._graph = None

    def quant_module_refactor(self, module: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        """Every Conv2d / Conv1d / Linear below `module` becomes a QuantModule."""
        for name, child in module.named_children():
            if isinstance(child, (nn.Conv2d, nn.Conv1d, nn.Linear)):
                setattr(module, name, QuantModule(child, weight_quant_params, act_quant_params))
            elif isinstance(child, StraightThrough):
                continue
            else:
                self.quant_module_refactor(child, weight_quant_params, act_quant_params)

    def quant_block_refactor(self, module: nn.Module, weight_quant_params: dict = {}, act_quant_params: dict = {}):
        for name, child in module.named_children():
            wrapper = self.specials.get(type(child))
            if wrapper is None:
                self.quant_block_refactor(child, weight_quant_params, act_quant_params)
            elif wrapper in (QuantBasicTransformerBlock, QuantAttnBlock):
                setattr(module, name, wrapper(child, act_quant_params, sm_abit=self.sm_abit))
            elif wrapper is QuantSMVMatMul:
                setattr(module, name, wrapper(act_quant_params, sm_abit=self.sm_abit))
            elif wrapper is QuantQKMatMul:
                setattr(module, name, wrapper(act_quant_params))
            else:
                setattr(module, name, wrapper(child, act_quant_params))
        if isinstance(getattr(module, 'qkv_matmul', None), QuantQKMatMul) and \
                isinstance(getattr(module, 'smv_matmul', None), QuantSMVMatMul) and hasattr(module, 'n_heads'):
            qattn.patch_legacy_attention(module)   # lets the two matmuls + softmax run as ONE fused kernel

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        for m in self.model.modules():
            if isinstance(m, (QuantModule, BaseQuantBlock)):
                m.set_quant_state(weight_quant, act_quant)

    def forward(self, x, timesteps=None, context=None):
        return self.model(x, timesteps, context)

    def set_grad_ckpt(self, grad_ckpt: bool):
        for _, m in self.model.named_modules():
            if isinstance(m, QuantBasicTransformerBlock) or type(m).__name__ == 'BasicTransformerBlock':
                m.checkpoint = grad_ckpt

    def set_first_last_layer_to_8bit(self):
        w_list, a_list = [], []
        for _, module in self.model.named_modules():
            if isinstance(module, UniformAffineQuantizer):
                (a_list if module.leaf_param else w_list).append(module)
        w_list[0].bitwidth_refactor(8)
        w_list[-1].bitwidth_refactor(8)
        a_list[-2].bitwidth_refactor(8)   # input of the last layer

    def disable_network_output_quantization(self):
        module_list = [m for m in self.model.modules() if isinstance(m, QuantModule)]
        module_list[-1].disable_act_quant = True

    # ---- B200 additions (not in the reference) -------------------------------------------------------
    def path_report(self):
        """{module name: 'int8' | 'fake' | 'fp'} taken by the last forward of each QuantModule."""
        return {n: m.last_path for n, m in self.model.named_modules() if isinstance(m, QuantModule)}
