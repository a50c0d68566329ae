"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the MACBF baseline path in plain torch fp32, on top of
the helpers of gcbf_oracle.py.  Only tests/ may import this module; the product package never does.

Follows reference gcbf/algo/macbf.py:20-239 (CBFNet, MACBF.update, MACBF.apply), gcbf/nn/gnn.py:82-135 (CBFNetLayer,
MACBFControllerLayer), gcbf/controller/macbf_controller.py:15-48 and the `max_neighbors` / `return_edge` branches of the three
environments.  Third-party arithmetic (PyG `MessagePassing(aggr='max')`, torch_cluster's `max_num_neighbors` cap) is restated
from the published algorithms: **parity unpinned** for those two pieces, like the GCBF oracle.  What pins this port:
tests/test_macbf_cpu.py runs the reference's own Python (oracle/ref_harness.py --algo macbf, on oracle/shim) in the build
container and compares edge lists, per-edge h, actions, masks, losses and post-step weights; tests/golden/macbf_*.pt hold the
reference-on-shim outputs for the GPU box.
"""
from typing import Dict, Optional

import numpy as np
import torch

import gcbf_oracle as O

Tensor = torch.Tensor
TOP_K = 12          # train.py:30, test.py:37: max_neighbors = 12 when --algo macbf

# gcbf/trainer/hyperparams.yaml, macbf rows
HYPERPARAMS = {
    'SimpleCar': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.0001, loss_unsafe_coef=1.0, loss_safe_coef=1.0,
                      loss_h_dot_coef=1.0),
    'SimpleDrone': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.01, loss_unsafe_coef=1.0, loss_safe_coef=1.0,
                        loss_h_dot_coef=1.0),
    'DubinsCar': dict(alpha=1.0, eps=0.02, inner_iter=10, loss_action_coef=0.0005, loss_unsafe_coef=1.0, loss_safe_coef=1.0,
                      loss_h_dot_coef=1.0),
}


def radius_graph_topk(env: str, pos: Tensor, num_agents: int, k: int = TOP_K) -> Tensor:
    """`add_communication_links` of an env built with max_neighbors = k, ONE graph.
    SimpleCar (simple_car.py:32-33 -> RadiusGraph(max_num_neighbors=k) -> torch_cluster, restated): the first k + 1 hits of
      `d2 < r*r` in ascending source index, the query itself included, then the self loop is dropped.
    DubinsCar / SimpleDrone (dubins_car.py:730-746, simple_drone.py:316-333): torch.topk(dist, k, largest=False) per agent row,
      every entry outside the top k is pushed beyond the radius, then `dist < r`."""
    r = O.ENV_PARAMS[env]['comm_radius']
    if env == 'SimpleCar':
        P = pos.detach().cpu().numpy().astype(np.float32)
        n = P.shape[0]
        d2 = np.zeros((n, n), dtype=np.float32)
        for d in range(P.shape[1]):
            diff = (P[:, None, d] - P[None, :, d]).astype(np.float32)
            d2 = (d2 + (diff * diff).astype(np.float32)).astype(np.float32)
        hit = d2 < np.float32(np.float32(r) * np.float32(r))
        hit &= np.cumsum(hit, axis=1) <= (k + 1)
        hit[np.arange(n), np.arange(n)] = False
        i, j = np.nonzero(hit)
        return torch.from_numpy(np.stack([j, i]).astype(np.int64))
    pos_diff = pos.unsqueeze(1) - pos.unsqueeze(0)
    dist = torch.norm(pos_diff, dim=-1)[:num_agents]
    dist[:, :num_agents] += torch.eye(num_agents) * (r + 1)
    _, dist_id = torch.topk(dist, k, dim=-1, largest=False)
    keep = torch.zeros_like(dist, dtype=torch.bool)
    keep.scatter_(1, dist_id, True)
    dist = torch.where(keep, dist, dist + (r + 1))
    return torch.nonzero(torch.less(dist, r), as_tuple=False).t()[[1, 0]]


def batch_radius_graph_topk(env: str, states: Tensor, num_graphs: int, nodes_per_graph: int, num_agents: int, k: int = TOP_K) -> Tensor:
    pd = O.ENV_PARAMS[env]['pos_dim']
    out = []
    for g in range(num_graphs):
        pos = states[g * nodes_per_graph:(g + 1) * nodes_per_graph, :pd]
        if env == 'SimpleCar':
            pos = pos[:num_agents]
        out.append(radius_graph_topk(env, pos, num_agents, k) + g * nodes_per_graph)
    return torch.cat(out, dim=1)


def edge_masks(env: str, e_attr: Tensor):
    """safe_mask / unsafe_mask(data, return_edge=True): simple_car.py:307-311, 332-336; dubins_car.py:819-823, 844-848;
    simple_drone.py:380-384, 405-409.  Returns (safe, unsafe) bool [E]."""
    p = O.ENV_PARAMS[env]
    dist = e_attr[:, :p['pos_dim']].norm(dim=-1)
    return torch.greater(dist, 4 * p['radius']), torch.less(dist, 2 * p['radius'])


def cbf_net_forward(sd: Dict[str, Tensor], x: Tensor, e_attr: Tensor, edge_index: Tensor) -> Tensor:
    """CBFNet.forward (macbf.py:32-51) = CBFNetLayer (gnn.py:82-104): h_ij = phi(cat[x_i, x_j, e_ij]) per EDGE, [E, 1]."""
    src, dst = edge_index[0], edge_index[1]
    info = torch.cat([x[dst], x[src], e_attr], dim=1)
    return O.mlp_forward(sd, 'net.module_0.phi', info, 4, False)


def controller_forward(sd: Dict[str, Tensor], x: Tensor, e_attr: Tensor, edge_index: Tensor, agent_mask: Optional[Tensor],
                       u_ref: Tensor) -> Tensor:
    """MACBFController.forward (macbf_controller.py:29-48) with MACBFControllerLayer (gnn.py:116-135): phi on the edges, MAX
    over the incoming messages (0 for a node without any), gamma, agent rows, cat with u_ref, head."""
    src, dst = edge_index[0], edge_index[1]
    info = torch.cat([x[dst], x[src], e_attr], dim=1)
    m = O.mlp_forward(sd, 'net.module_0.phi', info, 2, False)
    aggr = torch.zeros(x.shape[0], m.shape[1], dtype=m.dtype)
    aggr = aggr.scatter_reduce(0, dst.view(-1, 1).expand_as(m), m, reduce='amax', include_self=False)
    feat = O.mlp_forward(sd, 'net.module_0.gamma', aggr, 4, False)
    if agent_mask is not None:
        feat = feat[agent_mask]
    return O.mlp_forward(sd, 'feat_2_action', torch.cat([feat, u_ref], dim=1), 4, False)


def update_step(env: str, cbf_sd: Dict[str, Tensor], actor_sd: Dict[str, Tensor], opt_cbf: dict, opt_actor: dict, states: Tensor,
                goal: Tensor, edge_index: Tensor, u_ref_stored: Tensor, num_graphs: int, num_agents: int, num_obs: int,
                hp: Optional[dict] = None, K: Optional[Tensor] = None, apply_optim: bool = True, dt: float = O.DT) -> dict:
    """One inner iteration of MACBF.update (macbf.py:135-186) on a pre-collated batch; mutates the state dicts and Adam states."""
    hp = HYPERPARAMS[env] if hp is None else hp
    N = num_agents + num_obs
    x, agent_mask = O.make_graph_inputs(env, states, num_graphs, num_agents, num_obs)
    cbf_p = {k: cbf_sd[k].requires_grad_(True) for k in O.trainable_keys(cbf_sd)}
    act_p = {k: actor_sd[k].requires_grad_(True) for k in O.trainable_keys(actor_sd)}
    eps, alpha = hp['eps'], hp['alpha']
    e_attr = O.edge_attr(env, states, edge_index)
    h = cbf_net_forward(cbf_sd, x, e_attr, edge_index)                                     # macbf.py:137
    actions = controller_forward(actor_sd, x, e_attr, edge_index, agent_mask, u_ref_stored)  # macbf.py:138
    sm, um = edge_masks(env, e_attr)
    h_unsafe = h[um]                                                                       # macbf.py:144-153
    if h_unsafe.numel():
        loss_unsafe = torch.mean(torch.relu(h_unsafe + eps))
        acc_unsafe = torch.mean(torch.less(h_unsafe, 0).type_as(h_unsafe))
    else:
        loss_unsafe, acc_unsafe = torch.tensor(0.0), torch.tensor(1.0)
    h_safe = h[sm]                                                                         # macbf.py:156-164
    if h_safe.numel():
        loss_safe = torch.mean(torch.relu(-h_safe + eps))
        acc_safe = torch.mean(torch.greater_equal(h_safe, 0).type_as(h_safe))
    else:
        loss_safe, acc_safe = torch.tensor(0.0), torch.tensor(1.0)
    states_next = O.forward_states(env, states, agent_mask, actions, goal, K, N, dt)       # macbf.py:167 (retained edges)
    e_next = O.edge_attr(env, states_next, edge_index)
    h_next = cbf_net_forward(cbf_sd, x, e_next, edge_index)
    h_dot = (h_next - h) / dt
    loss_h_dot = torch.mean(torch.relu(-h_dot - alpha * h + eps))                          # macbf.py:169-170
    acc_h_dot = torch.mean(torch.greater_equal(h_dot + alpha * h, 0).type_as(h_dot))
    loss_action = torch.mean(torch.square(actions).sum(dim=1))                             # macbf.py:174
    loss = (hp['loss_unsafe_coef'] * loss_unsafe + hp['loss_safe_coef'] * loss_safe + hp['loss_h_dot_coef'] * loss_h_dot +
            hp['loss_action_coef'] * loss_action)
    plist = list(cbf_p.values()) + list(act_p.values())
    glist = torch.autograd.grad(loss, plist, allow_unused=True)
    gl = [g if g is not None else torch.zeros_like(p) for g, p in zip(glist, plist)]
    cbf_g = dict(zip(cbf_p.keys(), gl[:len(cbf_p)]))
    act_g = dict(zip(act_p.keys(), gl[len(cbf_p):]))
    for d in (cbf_sd, actor_sd):
        for k in d:
            d[k].requires_grad_(False)
    raw = dict(cbf={k: v.clone() for k, v in cbf_g.items()}, actor={k: v.clone() for k, v in act_g.items()})
    gn_cbf = O.clip_grad_norm(cbf_g, 1e-3)                                                 # macbf.py:183-184
    gn_act = O.clip_grad_norm(act_g, 1e-3)
    if apply_optim:
        with torch.no_grad():
            O.adam_step({k: cbf_sd[k] for k in cbf_g}, cbf_g, opt_cbf, lr=3e-4)            # macbf.py:84, 185
            O.adam_step({k: actor_sd[k] for k in act_g}, act_g, opt_actor, lr=1e-3)        # macbf.py:85, 186
    return dict(h=h.detach(), actions=actions.detach(), h_next=h_next.detach(), safe_mask=sm, unsafe_mask=um,
                states_next=states_next.detach(), loss_unsafe=loss_unsafe.detach(), loss_safe=loss_safe.detach(),
                loss_h_dot=loss_h_dot.detach(), loss_action=loss_action.detach(), loss=loss.detach(), acc_unsafe=acc_unsafe,
                acc_safe=acc_safe, acc_h_dot=acc_h_dot, grad_norm_cbf=gn_cbf, grad_norm_actor=gn_act, raw_grads=raw)


def apply_controller(env: str, cbf_sd: Dict[str, Tensor], actor_sd: Dict[str, Tensor], states: Tensor, goal: Tensor, edge_index: Tensor,
                     u_ref_stored: Tensor, num_agents: int, num_obs: int) -> Tensor:
    """MACBF.apply (macbf.py:209-239) on ONE graph.  The reference optimises `action = self.actor(data).detach()` with Adam(lr = 1),
    but that tensor is a leaf that does NOT require grad: `loss_h_dot.backward()` (which succeeds, because the CBF parameters do
    require grad) never gives it a gradient and `Adam.step()` skips parameters whose `.grad` is None.  The loop therefore only
    evaluates the CBF up to 32 times and the method returns the actor's action unchanged -- which is what this function returns
    (checked against the reference-on-shim: tests/test_macbf_cpu.py, fixture key `apply_action`)."""
    x, agent_mask = O.make_graph_inputs(env, states, 1, num_agents, num_obs)
    with torch.no_grad():
        e_attr = O.edge_attr(env, states, edge_index)
        return controller_forward(actor_sd, x, e_attr, edge_index, agent_mask, u_ref_stored)
