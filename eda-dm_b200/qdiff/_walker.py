"""Execution-order walk over a quantized UNet that hands every reconstruction unit to a callback.

The visiting order is the one the reference hard-codes (qdiff/recon_block_Qmodel.py:26-89): children in
definition order; the DDIM level `down.1` (the attention level) is unrolled block/attn/block/attn/downsample;
the `up` list is visited from its last level to its first, with `up.1` unrolled the same way.
"""
import logging

import torch.nn as nn

from .quant_layer import QuantModule
from .quant_block import BaseQuantBlock

logger = logging.getLogger(__name__)


class UnitWalker:
    def __init__(self, on_layer, on_block):
        self.on_layer, self.on_block = on_layer, on_block
        self._down_seen = None

    def _unit(self, name, module):
        if module.ignore_reconstruction is True:
            logger.info('Ignore reconstruction of %s', name)
            return
        if isinstance(module, QuantModule):
            logger.info('Reconstruction for layer %s', name)
            self.on_layer(module)
        else:
            logger.info('Reconstruction for block %s', name)
            self.on_block(module)

    def _unrolled_level(self, level, resample_attr):
        # the reference spells this out for 2 (down) / 3 (up) block+attention pairs; any count works the same way
        for i in range(len(level.block)):
            self.on_block(level.block[i])
            if i < len(level.attn):
                self.on_block(level.attn[i])
        if hasattr(level, resample_attr):
            self.on_layer(getattr(level, resample_attr).conv)

    def walk(self, parent: nn.Module):
        for name, module in parent.named_children():
            if self._down_seen is None and name == 'down':
                self._down_seen = 'down'
            if self._down_seen == 'down' and name == '1' and not isinstance(module, BaseQuantBlock):
                logger.info('reconstruction for down 1 modulelist')
                self._unrolled_level(module, 'downsample')
                self._down_seen = 'over'
            elif isinstance(module, (QuantModule, BaseQuantBlock)):
                self._unit(name, module)
            elif name == 'up':
                self.walk_up(module)
            else:
                self.walk(module)

    def walk_up(self, parent: nn.Module):
        for name, module in reversed(list(parent.named_children())):
            if name == '1':
                logger.info('reconstruction for up 1 modulelist')
                self._unrolled_level(module, 'upsample')
            elif isinstance(module, (QuantModule, BaseQuantBlock)):
                self._unit(name, module)
            else:
                self.walk(module)
