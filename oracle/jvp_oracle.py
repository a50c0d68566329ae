"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the analytic h_dot: h_dot = J_h(s) . f(s, u) with the edges held fixed,
computed by torch's autograd (double-backward JVP) through the GCBF port of gcbf_oracle.py.  Only tests/ may import this module.

There is no reference implementation of this quantity (the reference only forms the finite difference of gcbf/algo/gcbf.py:193-207),
so **parity unpinned** in the golden-vector sense; what pins it instead is the defining limit: tests/test_jvp_cpu.py checks this
oracle against a central finite difference of the same port in float64."""
import copy
from typing import Dict, Optional

import torch

import gcbf_oracle as O

Tensor = torch.Tensor


def closed_loop_state_dot(env: str, states: Tensor, goal: Tensor, action: Tensor, num_graphs: int, num_agents: int, num_obs: int,
                          K: Optional[Tensor] = None, freeze: Optional[bool] = None) -> Tensor:
    """f(x, clamp(action + u_ref(x))) for every node: the x_dot inside forward_graph (simple_car.py:178-194 and siblings,
    gcbf/env/base.py:381-398).  freeze defaults to the reference's single-graph discriminator (a batch of exactly one graph)."""
    p = O.ENV_PARAMS[env]
    x, am = O.make_graph_inputs(env, states, num_graphs, num_agents, num_obs)
    ag = states if am is None else states[am]
    tot = torch.clamp(action + O.u_ref(env, ag, goal, K), -p['action_lim'], p['action_lim'])
    if freeze is None:
        freeze = num_graphs == 1
    return O.dynamics(env, states, am, tot, goal, bool(freeze) and am is not None)


def cbf_of_states(env: str, cbf_sd: Dict[str, Tensor], states: Tensor, edge_index: Tensor, num_graphs: int, num_agents: int,
                  num_obs: int) -> Tensor:
    """h as a function of the states with the edge list fixed; every evaluation starts from the same spectral-norm buffers."""
    x, am = O.make_graph_inputs(env, states, num_graphs, num_agents, num_obs)
    sd = copy.deepcopy(cbf_sd)
    return O.cbf_forward(sd, x.to(states.dtype), O.edge_attr(env, states, edge_index), edge_index, am)


def h_and_h_dot(env: str, cbf_sd: Dict[str, Tensor], states: Tensor, goal: Tensor, edge_index: Tensor, action: Tensor, num_graphs: int,
                num_agents: int, num_obs: int, K: Optional[Tensor] = None, freeze: Optional[bool] = None):
    sdot = closed_loop_state_dot(env, states, goal, action, num_graphs, num_agents, num_obs, K, freeze)
    h, h_dot = torch.autograd.functional.jvp(lambda s: cbf_of_states(env, cbf_sd, s, edge_index, num_graphs, num_agents, num_obs),
                                             states, sdot)
    return h.detach(), h_dot.detach(), sdot
