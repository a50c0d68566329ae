"""Common base of the policy networks.  The contract is the reference's (gcbf/controller/base.py:8-48): a controller is an
`nn.Module` built from `(num_agents, node_dim, edge_dim, action_dim)` whose `forward(data)` maps a (possibly collated) graph
container to one action row per agent."""
from abc import ABC, abstractmethod
from typing import NamedTuple

import torch.nn as nn
from torch import Tensor


class _Dims(NamedTuple):
    num_agents: int
    node_dim: int
    edge_dim: int
    action_dim: int


class MultiAgentController(nn.Module, ABC):

    def __init__(self, num_agents: int, node_dim: int, edge_dim: int, action_dim: int):
        super().__init__()
        self._dims = _Dims(int(num_agents), int(node_dim), int(edge_dim), int(action_dim))

    num_agents = property(lambda self: self._dims.num_agents)
    node_dim = property(lambda self: self._dims.node_dim)
    edge_dim = property(lambda self: self._dims.edge_dim)
    action_dim = property(lambda self: self._dims.action_dim)

    def extra_repr(self) -> str:
        return ', '.join(f'{k}={v}' for k, v in self._dims._asdict().items())

    @abstractmethod
    def forward(self, data) -> Tensor:
        """data carries x, edge_attr, edge_index, u_ref [, agent_mask]; returns [num_graphs * num_agents, action_dim]."""
