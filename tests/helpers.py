"""Shared builders for the parity tests: the same (synthetic states, seeded weights) on the oracle side
(CPU tensors / state dicts) and on the product side (gcbf_b200 modules)."""
import copy

import torch

import gcbf_oracle as O
from gcbf_b200 import synth


def case_inputs(meta):
    sb = synth.make_states(meta['env'], meta['n'], meta['obs'], meta['graphs'], meta['area'], meta['seed'])
    return sb


def seeded_algo(env_name, n, device, init_seed=0, env_params=None, hyperparams='table'):
    """gcbf_b200 env + GCBF with the reference's seeded initialisation."""
    from gcbf_b200.algo import make_algo
    from gcbf_b200.env import make_env
    from gcbf_b200.trainer.utils import read_params
    env = make_env(env_name, n, device)
    params = env.default_params
    if env_params:
        params.update(env_params)
    env = make_env(env_name, n, device, params=params)
    torch.manual_seed(init_seed)
    hp = read_params(env_name, 'gcbf') if hyperparams == 'table' else hyperparams
    algo = make_algo('gcbf', env, n, env.node_dim, env.edge_dim, env.action_dim, device, 512, hp)
    return env, algo


def oracle_batch(sb):
    """edge_index / u_ref / x / agent_mask / K on the CPU oracle for a SynthBatch."""
    env, n, o, B = sb.env, sb.num_agents, sb.num_obs, sb.num_graphs
    N = n + o
    ei = O.batch_radius_graph(env, sb.states, B, N, n)
    K = O.lqr_gain(env) if env != 'DubinsCar' else None
    x, am = O.make_graph_inputs(env, sb.states, B, n, o)
    ag = sb.states if am is None else sb.states[am]
    ur = O.u_ref(env, ag, sb.goals, K)
    return dict(edge_index=ei, K=K, x=x, agent_mask=am, u_ref=ur, N=N)


def product_batch(env, sb, device):
    """gcbf_b200 graph for a SynthBatch: goal installed, radius graph + u_ref from the kernels."""
    env.set_goal(sb.goals)
    if sb.env == 'DubinsCar':
        env._obs = sb.obs.to(device)
    return env.graph_from_states(sb.states.to(device))


def sd_clone(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}
