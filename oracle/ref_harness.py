"""TEST INFRASTRUCTURE ONLY (build container only: needs /root/reference).

Runs the UNMODIFIED reference `gcbf` package (on oracle/shim) on a synthetic batch and returns every
intermediate the parity tests compare: edge_index, h, u, masks, the re-linked edge_index, the four losses,
three accuracies and the post-step state_dicts.  The only harness-side intervention is replacing
`Buffer.sample` by a function returning the prepared graph list (the reference samples the replay buffer with
host RNG, gcbf/algo/gcbf.py:151-156) and a dict-backed stand-in for the TensorBoard writer.

Usage as a script (used by tests/test_oracle_cpu.py via subprocess so that the product package and the
reference package never meet in one interpreter):
    python oracle/ref_harness.py --env DubinsCar --n 16 --obs 4 --graphs 3 --area 4.0 --seed 7 --out x.pt
"""
import argparse
import importlib.util
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _load_synth():
    spec = importlib.util.spec_from_file_location(
        'gcbf_b200_synth', os.path.join(ROOT, 'gcbf-pytorch_b200', 'gcbf_b200', 'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules['gcbf_b200_synth'] = mod
    spec.loader.exec_module(mod)
    return mod


class DictWriter:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step=None):
        self.scalars[tag] = float(value)


def build_reference(sb, init_seed=0, pretrained=None, hyperparams=None, algo_name='gcbf'):
    """Returns (env, algo, graph_list) with the synthetic state injected into the reference env."""
    sys.path.insert(0, HERE)
    from ref_loader import load_reference
    load_reference()
    from gcbf.env import make_env
    from gcbf.algo import make_algo
    from gcbf.trainer.utils import read_params
    from torch_geometric.data import Data
    from torch_geometric.utils import index_to_mask

    dev = torch.device('cpu')
    env = make_env(sb.env, sb.num_agents, dev)
    params = env.default_params
    params['area_size'] = sb.area_size
    params['num_obs'] = sb.num_obs
    max_nb = 12 if algo_name == 'macbf' else None          # train.py:30
    env = make_env(sb.env, sb.num_agents, dev, params=params, max_neighbors=max_nb)
    env.train()
    env._goal = sb.goals.clone()
    if sb.env != 'SimpleCar':
        env._obs = sb.obs.clone()
    n, N = sb.num_agents, sb.nodes_per_graph
    graphs = []
    for g in range(sb.num_graphs):
        st = sb.states[g * N:(g + 1) * N].clone()
        if sb.env == 'SimpleCar':
            d = Data(x=torch.zeros_like(st), pos=st[:, :2], states=st)
        else:
            pd = 2 if sb.env == 'DubinsCar' else 3
            d = Data(x=torch.cat((torch.zeros(n, 4), torch.ones(N - n, 4)), dim=0), pos=st[:, :pd], states=st,
                     agent_mask=index_to_mask(torch.arange(n), size=N))
        d = env.add_communication_links(d)
        d.update(Data(u_ref=env.u_ref(d)))
        graphs.append(d)
    torch.manual_seed(init_seed)
    hp = read_params(sb.env, algo_name) if hyperparams is None else hyperparams
    algo = make_algo(algo_name, env, n, env.node_dim, env.edge_dim, env.action_dim, dev, 512, hp)
    if pretrained is not None:
        algo.load(pretrained)
    return env, algo, graphs


def run_reference(sb, init_seed=0, pretrained=None, n_steps=1, algo_name='gcbf'):
    env, algo, graphs = build_reference(sb, init_seed, pretrained, algo_name=algo_name)
    edge = algo_name == 'macbf'                      # MACBF: per-edge h and per-edge masks
    from torch_geometric.data import Batch
    out = {}
    out['cbf_init'] = {k: v.clone() for k, v in algo.cbf.state_dict().items()}
    out['actor_init'] = {k: v.clone() for k, v in algo.actor.state_dict().items()}
    batch = Batch.from_data_list(graphs)
    out['edge_index'] = batch.edge_index.clone()
    out['u_ref'] = batch.u_ref.clone()
    out['edge_attr'] = batch.edge_attr.clone()
    # --- forward-only probes on *copies* (spectral norm mutates u/v on every forward) ----------------
    import copy
    cbf_c, actor_c = copy.deepcopy(algo.cbf), copy.deepcopy(algo.actor)
    with torch.no_grad():
        out['h_probe'] = cbf_c(batch).clone()
        out['u_probe'] = actor_c(batch).clone()
        out['unsafe_mask'] = env.unsafe_mask(batch, return_edge=edge).clone()
        out['safe_mask'] = env.safe_mask(batch, return_edge=edge).clone()
        nxt = env.forward_graph(batch, out['u_probe'])
        out['states_next_probe'] = nxt.states.clone()
    # --- test-time controller on the first graph (noise off: rand=0) ---------------------------------
    algo_c = copy.deepcopy(algo)
    algo_c._env = env
    out['apply_action'] = algo_c.apply(graphs[0], rand=0).detach().clone()
    # --- the real train step(s) ----------------------------------------------------------------------
    algo.params['inner_iter'] = 1
    algo.buffer.sample = lambda *a, **k: list(graphs)
    steps = []
    for it in range(n_steps):
        w = DictWriter()
        algo.buffer._data = list(graphs)            # update() clears the buffer after each call
        algo.memory._data = []
        res = algo.update(0, w)
        steps.append(dict(scalars=dict(w.scalars), ret=dict(res)))
    out['steps'] = steps
    out['cbf_final'] = {k: v.clone() for k, v in algo.cbf.state_dict().items()}
    out['actor_final'] = {k: v.clone() for k, v in algo.actor.state_dict().items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--env', required=True)
    ap.add_argument('--n', type=int, required=True)
    ap.add_argument('--obs', type=int, default=0)
    ap.add_argument('--graphs', type=int, default=1)
    ap.add_argument('--area', type=float, default=4.0)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--init-seed', type=int, default=0)
    ap.add_argument('--pretrained', default=None)
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--algo', default='gcbf', choices=['gcbf', 'macbf'])
    ap.add_argument('--out', required=True)
    a = ap.parse_args()
    synth = _load_synth()
    sb = synth.make_states(a.env, a.n, a.obs, a.graphs, a.area, a.seed)
    res = run_reference(sb, a.init_seed, a.pretrained, a.steps, a.algo)
    res['meta'] = dict(env=a.env, n=a.n, obs=sb.num_obs, graphs=a.graphs, area=a.area, seed=a.seed,
                       init_seed=a.init_seed, pretrained=a.pretrained, steps=a.steps, algo=a.algo)
    res['states'] = sb.states
    res['goals'] = sb.goals
    torch.save(res, a.out)


if __name__ == '__main__':
    main()
