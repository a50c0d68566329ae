"""Synthetic multi-agent state generator for the BASELINE.json configs.

The reference's `reset()` rejection sampling (gcbf/env/simple_car.py:100-106, dubins_car.py:405-418,
simple_drone.py:139-148) cannot terminate for n >~ 290 (2-D) at the default area, so benchmark and
parity inputs are drawn directly: positions i.i.d. U(0, area)^d (collisions therefore exist and the
unsafe mask is non-empty), per-env velocity/heading ranges as in section 8d.  Everything is drawn from a CPU
`torch.Generator` in fp32 so that the oracle, the golden fixtures and the CUDA path see identical bits.
"""
import math
from dataclasses import dataclass
import torch

ENV_DIMS = {  # state_dim, edge_dim, action_dim, pos_dim   (simple_car.py:43-57, dubins_car.py:102-108,
    'SimpleCar': (4, 4, 2, 2),   # simple_drone.py:47-61)
    'DubinsCar': (4, 5, 2, 2),
    'SimpleDrone': (6, 6, 3, 3),
}


@dataclass
class SynthBatch:
    env: str
    num_agents: int
    num_obs: int          # obstacles per graph actually present (SimpleCar: 0; SimpleDrone: num_agents)
    num_graphs: int
    area_size: float
    states: torch.Tensor  # [B*(n+o), state_dim]  agents first, then obstacles, per graph
    goals: torch.Tensor   # [n, goal_dim]  ONE goal set per env instance, shared by all graphs
    obs: torch.Tensor     # [o, state_dim] obstacle states of graph 0 (the env's `_obs`; shape matters)

    @property
    def nodes_per_graph(self) -> int:
        return self.num_agents + self.num_obs


def make_states(env: str, num_agents: int, num_obs: int, num_graphs: int, area_size: float,
                seed: int) -> SynthBatch:
    g = torch.Generator(device='cpu')
    g.manual_seed(seed)
    s_dim, _, _, p_dim = ENV_DIMS[env]
    n, B = num_agents, num_graphs

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    if env == 'SimpleCar':
        o = 0                                           # num_obs is ignored (simple_car.py:67-76)
        pos = U(B, n, 2, hi=area_size)
        vel = U(B, n, 2, lo=-0.5, hi=0.5)
        states = torch.cat([pos, vel], dim=2)
        goals = U(n, 2, hi=area_size)
        obs = torch.zeros(0, 4)
    elif env == 'DubinsCar':
        o = num_obs
        pos = U(B, n, 2, hi=area_size)
        theta = U(B, n, 1, lo=-math.pi, hi=math.pi)
        v = U(B, n, 1, hi=0.8)
        agents = torch.cat([pos, theta, v], dim=2)
        opos = U(B, o, 2, hi=area_size)                 # dubins_car.py:392-401
        otheta = U(B, o, 1, hi=2 * math.pi)
        ov = U(B, o, 1, hi=0.2)
        obstacles = torch.cat([opos, otheta, ov], dim=2)
        states = torch.cat([agents, obstacles], dim=1)
        goals = torch.cat([U(n, 2, hi=area_size), U(n, 1, lo=-math.pi, hi=math.pi),
                           torch.zeros(n, 1)], dim=1)   # dubins_car.py:443-446
        obs = obstacles[0].clone()
    elif env == 'SimpleDrone':
        o = n                                           # simple_drone.py:130-135: always n obstacles
        pos = U(B, n, 3, hi=area_size)
        vel = U(B, n, 3, lo=-0.3, hi=0.3)
        agents = torch.cat([pos, vel], dim=2)
        obstacles = torch.cat([U(B, o, 3, hi=area_size), torch.zeros(B, o, 3)], dim=2)
        states = torch.cat([agents, obstacles], dim=1)
        goals = torch.cat([U(n, 3, hi=area_size), torch.zeros(n, 3)], dim=1)
        obs = obstacles[0].clone()
    else:
        raise NotImplementedError(env)
    return SynthBatch(env, n, o, B, float(area_size), states.reshape(B * (n + o), s_dim).contiguous(),
                      goals.contiguous(), obs.contiguous())


# BASELINE.json configs -> concrete synthetic inputs.
CONFIGS = {
    'C1': dict(env='SimpleCar', num_agents=16, num_obs=0, num_graphs=1, area_size=4.0, seed=1001),
    'C2': dict(env='SimpleCar', num_agents=256, num_obs=8, num_graphs=32, area_size=16.0, seed=1002),
    'C3': dict(env='DubinsCar', num_agents=1024, num_obs=32, num_graphs=64, area_size=32.0, seed=1003),
    'C4': dict(env='SimpleDrone', num_agents=1024, num_obs=1024, num_graphs=16, area_size=8.0, seed=1004),
    'C5': dict(env='DubinsCar', num_agents=4096, num_obs=128, num_graphs=8, area_size=16.0, seed=1005),
}
# graphs one GPU owns (bench.py is weak scaling: every rank gets this many).  C4 / C5 are defined over 8 GPUs in
# BASELINE.json (16 replicas -> 2 per GPU, 8 dense graphs -> 1 per GPU); C1..C3 are single-GPU batches.
GRAPHS_PER_GPU = {'C1': 1, 'C2': 32, 'C3': 64, 'C4': 2, 'C5': 1, 'C1x256': 256}
# (C1x256: 256 copies of the reference's own training scale -- 16 agents in a 4 x 4 area -- as vectorised rollout environments)
CONFIGS['C1x256'] = dict(env='SimpleCar', num_agents=16, num_obs=0, num_graphs=256, area_size=4.0, seed=1011)


def seeded_algo(env_name, n, device, init_seed=0, env_params=None, hyperparams='table'):
    """gcbf_b200 env + GCBF with the reference's seeded initialisation: torch.manual_seed(init_seed) followed by the
    same construction order as reference gcbf/algo/gcbf.py:87-100 gives the same weights (up to LAPACK's QR in
    orthogonal_, which is not bit-reproducible across host CPUs)."""
    import torch
    from .algo import make_algo
    from .env import make_env
    from .trainer.utils import read_params
    env = make_env(env_name, n, device)
    params = env.default_params
    if env_params:
        params.update(env_params)
    env = make_env(env_name, n, device, params=params)
    torch.manual_seed(init_seed)
    hp = read_params(env_name, 'gcbf') if hyperparams == 'table' else hyperparams
    algo = make_algo('gcbf', env, n, env.node_dim, env.edge_dim, env.action_dim, device, 512, hp)
    return env, algo


def product_batch(env, sb, device):
    """gcbf_b200 graph for a SynthBatch: goal installed, radius graph + edge features + u_ref from the kernels."""
    env.set_goal(sb.goals)
    if sb.env == 'DubinsCar':
        env._obs = sb.obs.to(device)
    return env.graph_from_states(sb.states.to(device))
