from .base import MultiAgentController
from .gnn_controller import GNNController
