"""The MACBF baseline's policy network (reference gcbf/controller/macbf_controller.py:13-48): a max-aggregation message-passing
layer that maps every node to an `action_dim`-wide feature, restricted to the agent rows, concatenated with the nominal control
u_ref and reduced by a 4-layer head.  Same constructor, sub-module names (= checkpoint keys `net.module_0.{phi,gamma}.*`,
`feat_2_action.net.*`) and `forward(data)` contract as the reference."""
from torch import Tensor

from .. import ops
from ..data import agent_row_index
from ..nn.gnn import GraphSequential, MACBFControllerLayer
from ..nn.mlp import MLP
from .base import MultiAgentController


class MACBFController(MultiAgentController):
    HEAD_WIDTHS = (512, 128, 32)

    def __init__(self, num_agents: int, node_dim: int, edge_dim: int, phi_dim: int, action_dim: int):
        super().__init__(num_agents, node_dim, edge_dim, action_dim)
        self.net = GraphSequential(MACBFControllerLayer(node_dim=node_dim, edge_dim=edge_dim, output_dim=action_dim, phi_dim=phi_dim))
        self.feat_2_action = MLP(2 * action_dim, action_dim, self.HEAD_WIDTHS)

    def forward(self, data) -> Tensor:
        """data: x, edge_attr, edge_index, u_ref [, agent_mask]  ->  actions [num_graphs * num_agents, action_dim]."""
        feat = self.net(data.x, data.edge_attr, data.edge_index)
        return self.feat_2_action(ops.GatherCatFunction.apply(feat, agent_row_index(data), data.u_ref))
