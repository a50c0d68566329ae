"""The nominal baseline (reference gcbf/controller/nominal.py:9-22): the learned correction is zero, the env adds u_ref."""
import torch
from torch import Tensor

from ..data import agent_row_index
from .base import MultiAgentController


class NominalController(MultiAgentController):

    def forward(self, data) -> Tensor:
        rows = agent_row_index(data)
        n = int(rows.numel()) if rows is not None else int(data.states.shape[0])
        return torch.zeros(n, self.action_dim, device=data.states.device, dtype=data.states.dtype)
