"""Shim of torch_geometric.nn.aggr.AttentionalAggregation (PyG 2.3.0 aggr/attention.py):
gate = gate_nn(x); gate = softmax(gate, index, dim_size); out = scatter_sum(gate * x, index)."""
import torch
from ...utils import softmax


class AttentionalAggregation(torch.nn.Module):
    def __init__(self, gate_nn, nn=None):
        super().__init__()
        self.gate_nn = gate_nn
        self.nn = nn

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        gate = self.gate_nn(x)
        if self.nn is not None:
            x = self.nn(x)
        gate = softmax(gate, index, ptr, dim_size, dim=0)
        out = x.new_zeros((dim_size,) + tuple(x.shape[1:]))
        idx = index.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
        return out.scatter_add(0, idx, gate * x)
