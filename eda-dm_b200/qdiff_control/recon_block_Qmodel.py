"""recon_block_Qmodel, CFG variant (interface of the reference's qdiff_control/recon_block_Qmodel.py): plain
definition-order walk, every QuantModule a layer unit, every BaseQuantBlock a block unit."""
import logging

import torch.nn as nn

from qdiff.quant_layer import QuantModule
from qdiff.quant_block import BaseQuantBlock
from .block_recon import block_reconstruction
from .layer_recon import layer_reconstruction

logger = logging.getLogger(__name__)


class recon_block_Qmodel():
    def __init__(self, args, qnn, cali_data, kwargs):
        self.args, self.model, self.cali_data, self.kwargs = args, qnn, cali_data, kwargs
        self.down_name = None

    def recon_model(self, module: nn.Module):
        for name, child in module.named_children():
            if isinstance(child, (QuantModule, BaseQuantBlock)):
                if child.ignore_reconstruction is True:
                    logger.info('Ignore reconstruction of %s', name)
                    continue
                if isinstance(child, QuantModule):
                    logger.info('Reconstruction for layer %s', name)
                    layer_reconstruction(self.model, child, **self.kwargs)
                else:
                    logger.info('Reconstruction for block %s', name)
                    block_reconstruction(self.model, child, **self.kwargs)
            else:
                self.recon_model(child)

    def recon(self):
        self.recon_model(self.model)
        self.model.set_quant_state(weight_quant=True, act_quant=True)
        return self.model
