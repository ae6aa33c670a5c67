"""Forward-hook tap and seeding helpers (interface of the reference's qdiff/utils.py:12-54)."""
import os
import random

import numpy as np
import torch


class AttentionMap:
    """Keeps the last (input, output) of a module; FBR, TDAC and the cache builder read `.out` / `.feature`."""

    def __init__(self, module):
        self.hook = module.register_forward_hook(self.hook_fn)

    def hook_fn(self, module, input, output):
        self.out = output
        self.feature = input

    def remove(self):
        self.hook.remove()


def at(x):
    return x.view(x.size(0), -1)


def at_loss(x, y):
    return (at(x) - at(y)).pow(2).mean(1).sum()


def seed_everything(seed):
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    from .quant_layer import backend
    backend.qdrop_seed, backend.qdrop_offset = seed, 0
