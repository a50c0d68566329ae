from .mlp import MLP
from .gnn import ControllerGNNLayer, CBFGNNLayer, AttentionalAggregation, GraphSequential
