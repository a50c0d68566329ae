from .mlp import MLP
from .gnn import ControllerGNNLayer, CBFGNNLayer, CBFNetLayer, MACBFControllerLayer, AttentionalAggregation, GraphSequential
