"""Analytic h_dot: the directional derivative of the CBF along the closed-loop dynamics with the graph held fixed,

    h_dot_i = sum_k (dh_i / ds_k) . f(s_k, clamp(u_k + u_ref(s_k))),

as ONE forward-mode (tangent) pass next to the primal forward -- an additive alternative to the finite difference
(h(x + dt f) - h(x)) / dt the reference's loss uses (gcbf/algo/gcbf.py:193-207), meant for evaluation and diagnostics (the training
loss keeps the finite difference: parity with the reference, and differentiating a tangent pass would need second-order kernels).

The primal pass is the Python-sequenced GNN forward (ops.net_forward, which keeps every layer's activations); the tangent pass walks
the same layers: each linear layer is the SAME forward GEMM kernel applied to the tangent (no bias, no activation, the forward's
1/sigma), each activation multiplies by its derivative at the primal output (gcbf_act_bwd), the attention aggregation and the
two ends (state derivative, edge-feature tangent) have their own kernels (csrc/jvp.cu).  No torch arithmetic."""
import ctypes
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _C, ops
from ._C import call, ptr


def state_dot(env, data, action: Tensor, freeze: Optional[bool] = None) -> Tensor:
    """f(x, clamp(action + u_ref(x))) for every node of the batch: [B * N, state_dim].  freeze: the single-graph reach-freeze of
    the reference's dynamics(); default = what forward_graph does (a batch of exactly one graph)."""
    _C.require_cuda(data.states, action)
    B = env._num_graphs_of(data)
    freeze = (B == 1) if freeze is None else bool(freeze)
    st, ld = ops._mat(data.states.detach())
    act = action.detach().contiguous()
    uref = env.u_ref(data)
    goal_pg = getattr(data, 'goal', None) if hasattr(data, 'goal') else None
    goal, ldg = ops._mat((goal_pg if goal_pg is not None else env._goal).contiguous())
    out = torch.empty(st.shape[0], env.state_dim, device=st.device, dtype=torch.float32)
    cfg = env._cfg(B)
    call('gcbf_state_dot', ctypes.byref(cfg), ptr(st), ld, ptr(act), ptr(uref), ptr(goal), ldg, 1 if goal_pg is not None else 0,
         1 if freeze else 0, ptr(out), env.state_dim)
    return out


def edge_attr_tangent(env, states: Tensor, sdot: Tensor, edge_index: Tensor) -> Tensor:
    st, ld = ops._mat(states.detach())
    sd, ldsd = ops._mat(sdot)
    ei = edge_index.contiguous()
    E = int(ei.shape[1])
    out = torch.empty(E, env.edge_dim, device=st.device, dtype=torch.float32)
    call('gcbf_edge_attr_tangent', ops.ENV_IDS[env.ENV_NAME], ptr(st), ld, ptr(sd), ldsd, ptr(ei) if E else None, E, ptr(out) if E else None)
    return out


def mlp_tangent(ctx: ops.MLPCtx, layers, t: Tensor) -> Tensor:
    """Tangent of an MLP at the primal activations kept in `ctx` (ops.mlp_forward(..., save=True))."""
    if t.shape[0] == 0:                                   # a graph without edges: nothing to propagate through the edge MLPs
        return torch.empty(0, int(layers[-1].W.shape[0]), device=t.device, dtype=torch.float32)
    for l, L in enumerate(layers):
        N = int(L.W.shape[0])
        lin = ops.LinearSpec(L.W, torch.zeros(N, device=t.device, dtype=torch.float32), L.u, L.v, ops.ACT_NONE)
        t, _, _ = ops.mlp_forward(t, [lin], False, inv_sigmas=[ctx.inv_sigma[l]])       # the forward's sigma: no new power iteration
        if L.act != ops.ACT_NONE:
            t = ops.act_bwd(t, ctx.acts[l + 1], L.act)                                   # t * act'(y) from the primal output y
    return t


def net_tangent(spec: ops.NetSpec, ctx, t_edge_attr: Tensor, rowptr: Tensor, row_index: Optional[Tensor]) -> Tensor:
    """Tangent of ops.net_forward's output for a tangent of edge_attr (node features x are constants)."""
    c_phi, c_gate, c_gamma, c_head, msg, att, Nn, E = ctx
    dev = t_edge_attr.device
    C, nd, ed = spec.phi_dim, spec.node_dim, spec.edge_dim
    t_in = torch.zeros(E, 2 * nd + ed, device=dev, dtype=torch.float32)                 # d cat[x_i, x_j, e] = [0, 0, de]
    if E:
        ops.copy2d(t_edge_attr.contiguous(), t_in[:, 2 * nd:], E, ed)
    t_msg = mlp_tangent(c_phi, spec.phi, t_in)
    t_gate = mlp_tangent(c_gate, spec.gate, t_msg)
    t_gin_all = torch.zeros(Nn, C + nd, device=dev, dtype=torch.float32)                # d cat[aggr, x] = [d aggr, 0]
    call('gcbf_attn_aggr_tangent', ptr(msg) if E else None, C, ptr(t_msg) if E else None, C, ptr(att) if E else None,
         ptr(t_gate) if E else None, ptr(rowptr), Nn, C, ptr(t_gin_all), C + nd)
    if row_index is not None:
        t_gin = torch.empty(row_index.numel(), C + nd, device=dev, dtype=torch.float32)
        ops.rows_gather(t_gin_all, row_index, t_gin)
    else:
        t_gin = t_gin_all
    t = mlp_tangent(c_gamma, spec.gamma, t_gin)
    if spec.head is not None:
        t = mlp_tangent(c_head, spec.head, t)
    return t


def cbf_value_and_h_dot(cbf, env, data, action: Tensor, freeze: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
    """(h, h_dot) of a CBFGNN on a batch: h [B * n, 1] exactly as cbf(data) (one spectral-norm power iteration, like every forward of
    the reference), h_dot [B * n, 1] = dh/dt along x_dot = f(x, clamp(action + u_ref)) with the edges of `data` held fixed."""
    from .data import agent_row_index
    from .nn.gnn import cached_rowptr
    _C.require_cuda(data.states, data.edge_attr, data.edge_index, action)
    layer = cbf.feat_transformer.module_0
    spec = layer.net_spec(cbf.feat_2_CBF)
    x, ea, ei = data.x.contiguous(), data.edge_attr.detach().contiguous(), data.edge_index.contiguous()
    rowptr = cached_rowptr(data.edge_index, int(x.shape[0]))
    rows = agent_row_index(data)
    with torch.no_grad():
        h, ctx = ops.net_forward(spec, x, ea, ei, rowptr, rows, None, True)
        sdot = state_dot(env, data, action, freeze)
        t_ea = edge_attr_tangent(env, data.states, sdot, ei)
        h_dot = net_tangent(spec, ctx, t_ea, rowptr, rows)
    return h, h_dot
