"""Policy networks of the three algorithms `make_algo` knows: the GCBF actor (attention message passing, the north-star path), the
MACBF baseline's actor (max aggregation) and the nominal baseline (zero correction; the env adds u_ref)."""
from . import base, gnn_controller, macbf_controller, nominal

MultiAgentController = base.MultiAgentController
CONTROLLERS = {'gcbf': gnn_controller.GNNController, 'macbf': macbf_controller.MACBFController, 'nominal': nominal.NominalController}
GNNController, MACBFController, NominalController = CONTROLLERS['gcbf'], CONTROLLERS['macbf'], CONTROLLERS['nominal']
__all__ = ['MultiAgentController', 'CONTROLLERS', 'GNNController', 'MACBFController', 'NominalController']
