from typing import Optional

import torch

from .base import Algorithm
from .gcbf import GCBF, CBFGNN
from .macbf import MACBF, CBFNet
from .nominal import Nominal


def make_algo(algo: str, env, num_agents: int, node_dim: int, edge_dim: int, action_dim: int, device: torch.device,
              batch_size: int = 128, hyperparams: Optional[dict] = None) -> Algorithm:
    """Factory with the signature of reference gcbf/algo/__init__.py:12-36: 'gcbf' (the north-star hot path), 'macbf' (the paper's
    baseline; build the env with max_neighbors = 12 as train.py:30 does) and 'nominal'."""
    if algo == 'nominal':
        return Nominal(env, num_agents, node_dim, edge_dim, action_dim, device)
    if algo == 'gcbf':
        return GCBF(env, num_agents, node_dim, edge_dim, action_dim, device, batch_size, hyperparams)
    if algo == 'macbf':
        return MACBF(env, num_agents, node_dim, edge_dim, action_dim, device, batch_size, hyperparams)
    raise NotImplementedError('Unknown Algorithm!')
