"""Forward-hook tap and seeding helpers with the interface of the reference's qdiff/utils.py (AttentionMap :12-23,
at / at_loss :26-32, seed_everything :35-54); used by FBR, TDAC (scripts/calibration.py) and the cache builder."""
import os
import random

import numpy as np
import torch


class AttentionMap:
    """Tap on one module: after every forward, `.feature` is the tuple of positional inputs and `.out` the output."""

    def __init__(self, module):
        self.out = None
        self.feature = None
        self.hook = module.register_forward_hook(self.hook_fn)

    def hook_fn(self, module, input, output):
        self.feature, self.out = input, output

    def remove(self):
        if self.hook is not None:
            self.hook.remove()
            self.hook = None


def at(x):
    """flatten everything but the batch axis"""
    return x.reshape(x.shape[0], -1) if not x.is_contiguous() else x.view(x.shape[0], -1)


def at_loss(x, y):
    """batch sum of the per-sample mean squared difference"""
    diff = at(x) - at(y)
    return (diff * diff).mean(dim=1).sum()


def seed_everything(seed):
    """Seed every generator the reconstruction loop draws from (Python `random.sample`, numpy, torch CPU + CUDA) and the
    in-kernel QDrop stream; cuDNN is put in deterministic mode as the reference does."""
    os.environ["PYTHONHASHSEED"] = str(seed)
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    from .quant_layer import backend
    backend.qdrop_seed = seed
    backend.qdrop_offset = 0
