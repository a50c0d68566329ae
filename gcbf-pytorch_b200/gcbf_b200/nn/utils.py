"""Parameter initialisation shared by the MLP builders (the reference keeps the same helper in gcbf/nn/utils.py:4-7)."""
import torch
import torch.nn as nn


@torch.no_grad()
def init_param(module: nn.Module, gain: float = 1.) -> nn.Module:
    """Orthogonal weight (scaled by `gain`), zero bias, in place; returns the module.

    The draws must match the reference's seeded initialisation bit for bit (the golden fixtures and the pretrained
    checkpoints' sanity tests depend on it), so the weight goes through `nn.init.orthogonal_` on `weight.data` exactly once.
    For a spectral-normalised layer `module.weight` is the plain tensor aliasing `weight_orig`'s storage at construction time,
    so it is `weight_orig` that ends up orthogonal -- the reference's (accidental) behaviour, SURVEY 3.5, kept on purpose."""
    weight, bias = module.weight.data, module.bias.data
    nn.init.orthogonal_(weight, gain=gain)
    bias.zero_()
    return module
