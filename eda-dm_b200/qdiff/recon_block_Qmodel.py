"""recon_block_Qmodel / Change_LDM_model_attnblock (interface of the reference's qdiff/recon_block_Qmodel.py)."""
import logging

import torch.nn as nn

from .quant_block import QuantAttentionBlock, _ref_AttentionBlock
from .block_recon import block_reconstruction
from .layer_recon import layer_reconstruction
from ._walker import UnitWalker
from unet_zoo.ldm_unet import AttentionBlock as _ZooAttentionBlock

logger = logging.getLogger(__name__)

_ATTENTION_TYPES = tuple(c for c in (_ZooAttentionBlock, _ref_AttentionBlock) if c is not None)


def Change_LDM_model_attnblock(module: nn.Module, act_quant_params: dict = {}):
    """Wrap the remaining LDM AttentionBlocks so that they become reconstruction units."""
    for name, child in module.named_children():
        if isinstance(child, _ATTENTION_TYPES):
            setattr(module, name, QuantAttentionBlock(child, act_quant_params))
        else:
            Change_LDM_model_attnblock(child, act_quant_params)


class recon_block_Qmodel():
    def __init__(self, args, qnn, cali_data, kwargs):
        self.args, self.model, self.cali_data, self.kwargs = args, qnn, cali_data, kwargs
        self.down_name = None

    def recon_model(self, module: nn.Module):
        UnitWalker(lambda m: layer_reconstruction(self.model, m, **self.kwargs),
                   lambda m: block_reconstruction(self.model, m, **self.kwargs)).walk(module)

    def recon(self):
        self.recon_model(self.model)
        self.model.set_quant_state(weight_quant=True, act_quant=True)
        return self.model
