"""recon_layer_Qmodel (interface of the reference's qdiff/recon_layer_Qmodel.py): every QuantModule is its own
unit; a DDIM QuantAttnBlock is q, k, v, then its attention step sizes, then proj_out."""
import logging

import torch.nn as nn

from .quant_layer import QuantModule
from .quant_block import QuantResnetBlock, QuantAttnBlock
from .layer_recon import layer_reconstruction
from .attn_layer_recon import AttnBlock_layer_reconstruction
from ._walker import UnitWalker

logger = logging.getLogger(__name__)


class recon_layer_Qmodel():
    def __init__(self, args, qnn, cali_data, kwargs):
        self.args, self.model, self.cali_data, self.kwargs = args, qnn, cali_data, kwargs
        self.down_name = None

    def _layer(self, m):
        return layer_reconstruction(self.model, m, **self.kwargs)

    def recon_block(self, block: nn.Module):
        if isinstance(block, QuantResnetBlock):
            for m in block.modules():
                if isinstance(m, QuantModule) and m.ignore_reconstruction is not True:
                    self._layer(m)
        elif isinstance(block, QuantAttnBlock):
            for m in (block.q, block.k, block.v):
                self._layer(m)
            AttnBlock_layer_reconstruction(self.model, block, **self.kwargs)
            self._layer(block.proj_out)

    def recon_model(self, module: nn.Module):
        UnitWalker(self._layer, self.recon_block).walk(module)

    def recon(self):
        self.recon_model(self.model)
        self.model.set_quant_state(weight_quant=True, act_quant=True)
        return self.model
