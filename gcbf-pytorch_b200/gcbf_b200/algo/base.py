from abc import ABC, abstractmethod
from typing import Optional

import numpy as np
import torch
from torch import Tensor


class Algorithm(ABC):
    """Interface of reference gcbf/algo/base.py:13-189: act / step / post_step / is_update / update / save / load /
    apply, with `algo._env` re-assignable by the trainer (gcbf/trainer/trainer.py:91, 116)."""

    def __init__(self, env, num_agents: int, node_dim: int, edge_dim: int, action_dim: int, device: torch.device):
        self._env = env
        self._num_agents, self._node_dim, self._edge_dim, self._action_dim = num_agents, node_dim, edge_dim, action_dim
        self._device = device
        self.params = {}

    num_agents = property(lambda self: self._num_agents)
    node_dim = property(lambda self: self._node_dim)
    edge_dim = property(lambda self: self._edge_dim)
    action_dim = property(lambda self: self._action_dim)
    device = property(lambda self: self._device)

    @abstractmethod
    def act(self, data) -> Tensor:
        ...

    @abstractmethod
    def step(self, data, prob: float) -> Tensor:
        ...

    def post_step(self, data, action: Tensor, reward: float, done: bool, next_data):
        pass

    def sample(self, data, prob: float = 0.01) -> Tensor:
        actions = self.act(data)
        lo, hi = self._env.action_lim
        if np.random.uniform() < prob:
            actions = actions + torch.randn_like(actions) * 0.3 * (hi - lo)
        return actions

    @abstractmethod
    def is_update(self, step: int) -> bool:
        ...

    @abstractmethod
    def update(self, step: int, writer=None) -> dict:
        ...

    @abstractmethod
    def save(self, save_dir: str):
        ...

    @abstractmethod
    def load(self, load_dir: str):
        ...

    def apply(self, data, rand: Optional[float] = 30) -> Tensor:
        raise NotImplementedError
