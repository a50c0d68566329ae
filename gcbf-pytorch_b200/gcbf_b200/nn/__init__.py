"""Network building blocks with the reference's names (gcbf/nn): `MLP`, the two attention layers of GCBF, the two MACBF layers, plus the
stand-ins for the two torch_geometric containers whose attribute names end up in checkpoint keys."""
from .gnn import (AttentionalAggregation, CBFGNNLayer, CBFNetLayer, ControllerGNNLayer, GraphSequential,
                  MACBFControllerLayer)
from .mlp import MLP

__all__ = ['MLP', 'CBFGNNLayer', 'ControllerGNNLayer', 'CBFNetLayer', 'MACBFControllerLayer', 'AttentionalAggregation', 'GraphSequential']
