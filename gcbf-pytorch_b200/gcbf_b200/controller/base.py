from abc import ABC, abstractmethod

import torch.nn as nn
from torch import Tensor


class MultiAgentController(nn.Module, ABC):
    """Interface of reference gcbf/controller/base.py:8-48."""

    def __init__(self, num_agents: int, node_dim: int, edge_dim: int, action_dim: int):
        super().__init__()
        self._num_agents, self._node_dim, self._edge_dim, self._action_dim = num_agents, node_dim, edge_dim, action_dim

    num_agents = property(lambda self: self._num_agents)
    node_dim = property(lambda self: self._node_dim)
    edge_dim = property(lambda self: self._edge_dim)
    action_dim = property(lambda self: self._action_dim)

    @abstractmethod
    def forward(self, data) -> Tensor:
        """data: graph container with x, edge_attr, edge_index, u_ref[, agent_mask] -> (B*n, action_dim)."""
