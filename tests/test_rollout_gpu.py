"""Vectorised device rollouts (gcbf_b200/algo/rollout.py): B independent environments -- own states, own goal sets --
advance as one batch.  Per env the step must be the reference's single-env step (gcbf/env/*.py `step`: u_ref from the env's goal,
clamp, dynamics with the single-graph reach-freeze branch, gcbf/algo/gcbf.py:128-139 actor forward): checked against the oracle port
env by env, and against this package's own single-env path."""
import numpy as np
import pytest
import torch

import gcbf_oracle as O
from gcbf_b200 import synth
from gcbf_b200.algo.rollout import VectorRollout
from helpers import sd_clone, seeded_algo

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0') if torch.cuda.is_available() else None


def _goals(env_name, n, B, area, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(B):
        sb = synth.make_states(env_name, n, 0 if env_name == 'SimpleCar' else 4, 1, area, int(torch.randint(0, 10 ** 6, (1,), generator=g)))
        out.append(sb.goals)
    return torch.cat(out, dim=0)


@pytest.mark.parametrize('env_name,n,obs,B,area', [('DubinsCar', 16, 4, 5, 2.0), ('SimpleCar', 12, 0, 6, 1.5), ('SimpleDrone', 8, 8, 4, 1.0)])
def test_vector_step_is_the_single_env_step_per_env(env_name, n, obs, B, area):
    sb = synth.make_states(env_name, n, obs, B, area, 201)
    env, algo = seeded_algo(env_name, n, DEV, 0, {'num_obs': sb.num_obs, 'area_size': area})
    if env_name == 'DubinsCar':
        env._obs = sb.obs.to(DEV)
    goals = _goals(env_name, n, B, area, 7)
    # put two agents of env 1 on / next to their goals: the reach-freeze branch must trigger for that env only
    N = sb.nodes_per_graph
    gd = goals.shape[1]
    pd = env.POS_DIM
    sb.states[N + 0, :pd] = goals[n + 0, :pd]
    sb.states[N + 2, :pd] = goals[n + 2, :pd] + 0.004
    env.set_goal(goals[:n])
    vr = VectorRollout(env, algo, B, states=sb.states, goals=goals)
    before = vr.states.clone()
    out = vr.step(prob=0.0, store=False)
    nxt = vr.states.cpu()
    act = sd_clone(algo.actor)
    K = O.lqr_gain(env_name) if env_name != 'DubinsCar' else None
    for i in range(B):
        st = sb.states[i * N:(i + 1) * N]
        gl = goals[i * n:(i + 1) * n]
        ei = O.radius_graph(env_name, st[:, :pd] if env_name != 'SimpleCar' else st[:n, :pd], n)
        x, am = O.make_graph_inputs(env_name, st, 1, n, sb.num_obs)
        ur = O.u_ref(env_name, st if am is None else st[am], gl, K)
        with torch.no_grad():
            u = O.actor_forward(act, x, O.edge_attr(env_name, st, ei), ei, am, ur)
        assert (out['action'][i * n:(i + 1) * n].cpu() - u).abs().max().item() <= 1e-5
        want = O.forward_states(env_name, st, am, u, gl, K, N)                      # single graph => freeze branch
        assert (nxt[i * N:(i + 1) * N] - want).abs().max().item() <= 2e-6, i
        reach = (want[:n, :pd] - gl[:, :pd]).norm(dim=1) < O.ENV_PARAMS[env_name]['dist2goal']
        assert torch.equal(out['reach'][i].cpu(), reach)
    # the frozen agents of env 1 did not move (Dubins / Drone); SimpleCar has no freeze branch
    if env_name != 'SimpleCar':
        assert torch.equal(nxt[N + 0], before[N + 0].cpu())


def test_rollout_feeds_the_replay_ring_and_the_train_step_uses_per_graph_goals():
    env_name, n, obs, B, area = 'DubinsCar', 16, 4, 6, 2.5
    sb = synth.make_states(env_name, n, obs, B, area, 202)
    env, algo = seeded_algo(env_name, n, DEV, 0, {'num_obs': sb.num_obs, 'area_size': area})
    env._obs = sb.obs.to(DEV)
    goals = _goals(env_name, n, B, area, 8)
    env.set_goal(goals[:n])
    algo.use_device_replay(capacity=64)
    vr = VectorRollout(env, algo, B, states=sb.states, goals=goals)
    np.random.seed(0)
    for _ in range(4):
        vr.step(prob=0.3)
    assert algo.buffer.size == 4 * B
    algo.batch_size = 40
    algo.params['inner_iter'] = 1
    info = algo.update(1, None)                               # samples the ring (resolving the safe / unsafe flags), collates, trains
    assert set(info) == {'acc/safe', 'acc/unsafe', 'acc/derivative'} and all(v == v for v in info.values())
    assert len(algo.memory.safe_data) + len(algo.memory.unsafe_data) == 4 * B
    # per-graph goals change the step: the same batch with ONE shared goal set gives a different h_next
    from gcbf_b200.algo.device_buffer import collate
    idx = list(range(8))
    batch = collate(env, [(algo.memory, idx)])
    assert hasattr(batch, 'goal') and batch.goal.shape == (8 * n, goals.shape[1])
    res_pg = algo.train_step(batch, apply_optim=False)
    hn_pg = res_pg['h_next'].clone()
    from gcbf_b200.data import Data
    shared = Data(**{k: batch[k] for k in batch.keys() if k != 'goal'})
    res_sh = algo.train_step(shared, apply_optim=False)
    assert (res_sh['h_next'] - hn_pg).abs().max().item() > 0


@pytest.mark.parametrize('env_name,n,obs,area', [('DubinsCar', 12, 4, 2.0), ('SimpleCar', 14, 0, 1.5), ('SimpleDrone', 6, 6, 1.0)])
def test_device_collate_against_the_oracle(env_name, n, obs, area):
    """`Batch.from_data_list` (gcbf/algo/gcbf.py:159: block-diagonal collation with edge_index offsets) as the device ring does it -- two
    gathers + ONE batched radius graph + edge features -- against the ORACLE's collation of the same graphs (per-graph radius graph +
    offsets, edge features, node types): edge_index bit-exact, everything else equal."""
    from gcbf_b200.algo.device_buffer import DeviceReplay, collate
    env, algo = seeded_algo(env_name, n, DEV, 0, {'num_obs': obs, 'area_size': area})
    ring = DeviceReplay(DEV, capacity=8)
    stored = []
    goal = None
    for k in range(11):                                                      # wraps the ring's initial capacity
        sbk = synth.make_states(env_name, n, obs, 1, area, 700 + k)
        if goal is None:
            goal = sbk.goals
            env.set_goal(goal)
            if env_name == 'DubinsCar':
                env._obs = sbk.obs.to(DEV)
        g = env.graph_from_states(sbk.states.to(DEV))
        ring.append(g, is_safe=(k % 2 == 0))
        stored.append(sbk)
    idx = [9, 2, 3, 10, 0]
    batch = collate(env, [(ring, idx)])
    N = stored[0].nodes_per_graph
    states = torch.cat([stored[i].states for i in idx], dim=0)
    ei = O.batch_radius_graph(env_name, states, len(idx), N, n)
    assert torch.equal(batch.edge_index.cpu(), ei)
    assert torch.equal(batch.states.cpu(), states)
    assert (batch.edge_attr.cpu() - O.edge_attr(env_name, states, ei)).abs().max().item() <= 1e-6
    x, am = O.make_graph_inputs(env_name, states, len(idx), n, stored[0].num_obs)
    assert torch.equal(batch.x.cpu(), x)
    if am is not None:
        assert torch.equal(batch.agent_mask.cpu(), am)
    K = O.lqr_gain(env_name) if env_name != 'DubinsCar' else None
    ur = O.u_ref(env_name, states if am is None else states[am], goal, K)
    assert (batch.u_ref.cpu() - ur).abs().max().item() <= 1e-5
