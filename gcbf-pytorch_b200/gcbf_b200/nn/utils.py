import torch.nn as nn


def init_param(module: nn.Module, gain: float = 1.):
    """Orthogonal weight / zero bias (reference gcbf/nn/utils.py:4-7).  For a spectral-normalised layer
    `module.weight` is the plain tensor that aliases `weight_orig`'s storage, so `weight_orig` is what
    ends up orthogonal -- the same (accidental) behaviour as the reference."""
    nn.init.orthogonal_(module.weight.data, gain=gain)
    nn.init.constant_(module.bias.data, 0)
    return module
