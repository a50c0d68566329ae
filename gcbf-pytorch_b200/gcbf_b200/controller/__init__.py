from .base import MultiAgentController
from .nominal import NominalController
from .gnn_controller import GNNController
from .macbf_controller import MACBFController
