from typing import Optional

import torch

from .base import MultiAgentEnv
from .simple_car import SimpleCar
from .dubins_car import DubinsCar
from .simple_drone import SimpleDrone

_ENVS = {'SimpleCar': SimpleCar, 'SimpleDrone': SimpleDrone, 'DubinsCar': DubinsCar}


def make_env(env: str, num_agents: int, device: torch.device, dt: float = 0.03, params: Optional[dict] = None,
             max_neighbors: Optional[int] = None) -> MultiAgentEnv:
    """Factory with the signature of reference gcbf/env/__init__.py:11-26."""
    if env not in _ENVS:
        raise NotImplementedError('Env name not supported!')
    return _ENVS[env](num_agents, device, dt, params, max_neighbors)
