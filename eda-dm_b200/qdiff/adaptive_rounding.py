"""AdaRoundQuantizer with the reference's interface (qdiff/adaptive_rounding.py:9-78) over the fused
AdaRound kernels (edadm_adaround_fwd / _bwd / _init_alpha)."""
import logging

import torch
from torch import nn

from edadm import ops
from .quant_layer import UniformAffineQuantizer, round_ste

logger = logging.getLogger(__name__)


class AdaRoundQuantizer(nn.Module):
    """Learned rounding: W~ = (clamp(floor(W/d) + h(alpha) + zp, 0, L-1) - zp) * d with
    h = clamp(sigmoid(alpha)*(zeta-gamma)+gamma, 0, 1) while `soft_targets`, (alpha >= 0) afterwards."""

    def __init__(self, uaq: UniformAffineQuantizer, weight_tensor: torch.Tensor, round_mode='learned_round_sigmoid'):
        super().__init__()
        self.n_bits = uaq.n_bits
        self.sym = uaq.sym
        self.delta = uaq.delta
        self.zero_point = uaq.zero_point
        self.n_levels = uaq.n_levels
        self.round_mode = round_mode
        self.alpha = None
        self.soft_targets = False
        self.gamma, self.zeta = -0.1, 1.1
        self.beta = 2 / 3
        self.init_alpha(x=weight_tensor.clone())

    def forward(self, x):
        if self.round_mode == 'learned_hard_sigmoid':
            return ops.adaround_fake_quant(x, self.alpha, self.delta, self.zero_point, self.n_levels, self.soft_targets)
        if self.round_mode == 'nearest':
            return ops.uaq_forward(x, self.delta, self.zero_point, self.n_levels)
        if self.round_mode == 'nearest_ste':
            return ops.uaq_fake_quant(x, self.delta, self.zero_point, self.n_levels)
        if self.round_mode == 'stochastic':
            x_floor = torch.floor(x / self.delta)
            x_int = x_floor + torch.bernoulli((x / self.delta) - x_floor)
            logger.info('Draw stochastic sample')
            x_quant = torch.clamp(x_int + self.zero_point, 0, self.n_levels - 1)
            return (x_quant - self.zero_point) * self.delta
        raise ValueError('Wrong rounding mode')

    def get_soft_targets(self):
        return torch.clamp(torch.sigmoid(self.alpha) * (self.zeta - self.gamma) + self.gamma, 0, 1)

    def codes(self, x):
        """uint8 codes of the hard rounding (what the integer path packs)."""
        return ops.adaround_forward(x, self.alpha, self.delta, self.zero_point, self.n_levels, False, want_codes=True)[1]

    def init_alpha(self, x: torch.Tensor):
        if self.round_mode != 'learned_hard_sigmoid':
            raise NotImplementedError
        self.alpha = nn.Parameter(ops.adaround_init_alpha(x, self.delta))

    def extra_repr(self):
        return 'bit={n_bits}, symmetric={sym}, round_mode={round_mode}'.format(**self.__dict__)
