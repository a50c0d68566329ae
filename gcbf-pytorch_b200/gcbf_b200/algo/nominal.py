"""The nominal baseline (reference gcbf/algo/nominal.py:14-59): no learning, the action correction is zero (the env adds u_ref)."""
from typing import Optional

import torch
from torch import Tensor

from ..controller import NominalController
from .base import Algorithm


class Nominal(Algorithm):

    def __init__(self, env, num_agents: int, node_dim: int, edge_dim: int, action_dim: int, device: torch.device):
        super().__init__(env=env, num_agents=num_agents, node_dim=node_dim, edge_dim=edge_dim, action_dim=action_dim, device=device)
        self.actor = NominalController(num_agents=num_agents, node_dim=node_dim, edge_dim=edge_dim, action_dim=action_dim).to(device)

    def step(self, data, prob: float) -> Tensor:
        raise NotImplementedError

    def is_update(self, step: int) -> bool:
        raise NotImplementedError

    def update(self, step: int, writer=None):
        raise NotImplementedError

    def save(self, save_dir: str):
        raise NotImplementedError

    def load(self, load_dir: str):
        raise NotImplementedError

    def act(self, data) -> Tensor:
        with torch.no_grad():
            return self.actor(data)

    def apply(self, data, rand: Optional[float] = 30) -> Tensor:
        return self.act(data)
