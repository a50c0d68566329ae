from typing import Optional

import torch

from .base import Algorithm
from .gcbf import GCBF, CBFGNN


def make_algo(algo: str, env, num_agents: int, node_dim: int, edge_dim: int, action_dim: int, device: torch.device,
              batch_size: int = 128, hyperparams: Optional[dict] = None) -> Algorithm:
    """Factory with the signature of reference gcbf/algo/__init__.py:12-36.  Only 'gcbf' is on the hot path this
    package implements; 'macbf' / 'nominal' are the paper's baselines."""
    if algo == 'gcbf':
        return GCBF(env, num_agents, node_dim, edge_dim, action_dim, device, batch_size, hyperparams)
    raise NotImplementedError(f"algorithm {algo!r}: only 'gcbf' is implemented by gcbf_b200 (MACBF / nominal are "
                              'outside the north-star hot path)')
