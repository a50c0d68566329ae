"""GNNController (reference gcbf/controller/gnn_controller.py:13-48): ControllerGNNLayer(out=1024) ->
agent rows -> MLP(1024 + action_dim -> 512,128,32 -> action_dim) on cat[feat, u_ref], as ONE fused
autograd node (ops.GNNNetFunction)."""
from torch import Tensor

from ..data import agent_row_index
from ..nn.gnn import ControllerGNNLayer, GraphSequential
from ..nn.mlp import MLP
from .base import MultiAgentController


class GNNController(MultiAgentController):

    def __init__(self, num_agents: int, node_dim: int, edge_dim: int, phi_dim: int, action_dim: int):
        super().__init__(num_agents=num_agents, node_dim=node_dim, edge_dim=edge_dim, action_dim=action_dim)
        self.feat_transformer = GraphSequential(
            ControllerGNNLayer(node_dim=node_dim, edge_dim=edge_dim, output_dim=1024, phi_dim=phi_dim))
        self.feat_2_action = MLP(in_channels=1024 + action_dim, out_channels=action_dim, hidden_layers=(512, 128, 32))

    def forward(self, data) -> Tensor:
        layer = self.feat_transformer.module_0
        return layer.run(data.x, data.edge_attr, data.edge_index, row_index=agent_row_index(data),
                         head=self.feat_2_action, head_extra=data.u_ref)
