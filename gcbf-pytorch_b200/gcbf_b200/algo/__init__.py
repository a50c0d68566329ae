"""Algorithms: GCBF (the north-star hot path), the paper's MACBF baseline and the nominal controller."""
from typing import Optional

import torch

from .base import Algorithm
from .gcbf import CBFGNN, GCBF
from .macbf import MACBF, CBFNet
from .nominal import Nominal

_TRAINABLE = {'gcbf': GCBF, 'macbf': MACBF}


def make_algo(algo: str, env, num_agents: int, node_dim: int, edge_dim: int, action_dim: int, device: torch.device,
              batch_size: int = 128, hyperparams: Optional[dict] = None) -> Algorithm:
    """Factory with the signature of reference gcbf/algo/__init__.py:12-36.  For 'macbf' build the env with max_neighbors = 12 first,
    as the reference's train.py:30 does."""
    dims = (env, num_agents, node_dim, edge_dim, action_dim, device)
    if algo in _TRAINABLE:
        return _TRAINABLE[algo](*dims, batch_size, hyperparams)
    if algo == 'nominal':
        return Nominal(*dims)
    raise NotImplementedError('Unknown Algorithm!')
