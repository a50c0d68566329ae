"""End-to-end parity on the GPU, through the public Python API (which calls the C ABI):
  * every committed golden fixture (outputs of the UNMODIFIED reference, tests/golden/*.pt): edge_index bit-exact;
    h, u, losses within 1e-5 (the tolerance BASELINE.json's north_star states, fp32); masks equal; post-step
    weights within 1e-5 after two train steps;
  * the CPU oracle run live on fresh seeds;
  * size-independent properties at the BASELINE C2 size (oracle too slow there for the whole step).
"""
import pytest
import torch

import gcbf_oracle as O
from conftest import digest_close, golden_cases, load_golden
from gcbf_b200 import ops, synth
from gcbf_b200.data import Batch, Data
from helpers import case_inputs, oracle_batch, product_batch, sd_clone, seeded_algo

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0') if torch.cuda.is_available() else None
TOL = 1e-5   # north_star: h, u, loss within 1e-5 fp32


def _prepare(meta, case=''):
    sb = case_inputs(meta)
    if case.endswith('freeze'):
        sb.states[0, :2] = sb.goals[0, :2]
        sb.states[3, :2] = sb.goals[3, :2] + 0.01
    env, algo = seeded_algo(meta['env'], meta['n'], DEV, meta.get('init_seed', 0),
                            {'num_obs': sb.num_obs, 'area_size': sb.area_size})
    data = product_batch(env, sb, DEV)
    return sb, env, algo, data


@pytest.mark.parametrize('case', golden_cases())
def test_golden_forward(case):
    fix = load_golden(case)
    sb, env, algo, data = _prepare(fix['meta'], case)
    assert torch.equal(data.edge_index.cpu(), fix['edge_index'])                               # bit-exact
    assert torch.allclose(data.u_ref.cpu(), fix['u_ref'], rtol=0, atol=TOL)
    assert torch.allclose(data.edge_attr.cpu(), fix['edge_attr'], rtol=0, atol=TOL)
    # seeded init: orthogonal_() goes through LAPACK QR, which is not bit-reproducible across host CPUs
    assert not digest_close(sd_clone(algo.cbf), fix['cbf_init'], 1e-6, 1e-6)
    with torch.no_grad():
        h = algo.cbf(data)
        u = algo.actor(data)
    assert h.shape == fix['h_probe'].shape and u.shape == fix['u_probe'].shape
    assert (h.cpu() - fix['h_probe']).abs().max().item() <= TOL
    assert (u.cpu() - fix['u_probe']).abs().max().item() <= TOL
    assert torch.equal(env.unsafe_mask(data).cpu(), fix['unsafe_mask'])
    assert torch.equal(env.safe_mask(data).cpu(), fix['safe_mask'])
    nxt = env.forward_graph(data, fix['u_probe'].to(DEV))
    assert torch.allclose(nxt.states.cpu(), fix['states_next_probe'], rtol=0, atol=2e-6)


@pytest.mark.parametrize('case', golden_cases())
def test_golden_train_steps(case):
    fix = load_golden(case)
    sb, env, algo, data = _prepare(fix['meta'], case)
    for gold in fix['steps']:
        res = algo.train_step(data)
        s = res['scalars'].tolist()
        got = {'loss/unsafe': s[0], 'loss/safe': s[1], 'loss/derivative': s[2], 'loss/action': s[3],
               'acc/unsafe': s[4], 'acc/safe': s[5], 'acc/derivative': float(res['acc_h_dot'])}
        for tag, want in gold['scalars'].items():
            tol = TOL if tag.startswith('loss') else 1.5 / max(1, int(s[7]))     # accuracies: one flipped sample
            assert abs(got[tag] - want) <= tol, (tag, got[tag], want)
    # post-step weights after two clipped Adam steps (lr 1e-3 / 3e-4): digests within 1e-5 plus a budget of
    # 1% of the elements taking a sign-flipped first Adam step (see conftest.digest_close)
    bad = digest_close(sd_clone(algo.actor), fix['actor_final'], 1e-5, 1e-5, flip=0.01 * 2 * 1e-3 * 2)
    assert not bad, bad[:3]
    bad = digest_close(sd_clone(algo.cbf), fix['cbf_final'], 1e-5, 1e-5, flip=0.01 * 2 * 3e-4 * 2)
    assert not bad, bad[:3]


@pytest.mark.parametrize('env_name,n,obs,B,area,seed', [('SimpleCar', 32, 0, 4, 3.0, 41), ('DubinsCar', 32, 5, 3, 3.0, 42),
                                                        ('SimpleDrone', 16, 16, 3, 1.2, 43)])
def test_train_step_against_live_oracle(env_name, n, obs, B, area, seed):
    meta = dict(env=env_name, n=n, obs=obs, graphs=B, area=area, seed=seed, init_seed=1)
    sb, env, algo, data = _prepare(meta)
    cbf, act = sd_clone(algo.cbf), sd_clone(algo.actor)
    ob = oracle_batch(sb)
    want = O.update_step(env_name, cbf, act, {}, {}, sb.states, sb.goals, ob['edge_index'], ob['u_ref'], B, n, sb.num_obs,
                         K=ob['K'])
    res = algo.train_step(data)
    assert (res['h'].cpu() - want['h']).abs().max().item() <= TOL
    assert (res['actions'].cpu() - want['actions']).abs().max().item() <= TOL
    assert (res['h_next'].cpu().reshape(-1) - want['h_next']).abs().max().item() <= TOL
    assert torch.equal(res['edge_index_new'].cpu(), want['edge_index_new'])
    assert (res['h_next_new'].cpu().reshape(-1) - want['h_next_new']).abs().max().item() <= TOL
    s = res['scalars'].tolist()
    for got, key in zip(s[:4], ('loss_unsafe', 'loss_safe', 'loss_h_dot', 'loss_action')):
        assert abs(got - float(want[key])) <= TOL, (key, got, float(want[key]))
    # raw (pre-clip) gradients, relative to each tensor's scale
    # weights after one clipped Adam step: all but a small fraction of elements (sign-like first Adam step on
    # noise-level gradients: update = lr * g / (|g| + 1e-8) with |g| ~ 1e-9 after the 1e-3 norm clip) agree to 2e-5;
    # nothing may move by more than 2*lr
    for mod, ref_sd, lr in ((algo.cbf, cbf, 3e-4), (algo.actor, act, 1e-3)):
        for k, v in mod.state_dict().items():
            diff = (v.cpu() - ref_sd[k]).abs()
            tol = 2e-5 + 1e-4 * ref_sd[k].abs()
            assert (diff > tol).float().mean().item() <= 0.10, (k, (diff > tol).float().mean().item())
            if k.endswith(('weight', 'bias', 'weight_orig')):
                assert diff.max().item() <= 2 * lr + 1e-6, (k, diff.max().item())


def test_raw_gradients_against_live_oracle():
    meta = dict(env='DubinsCar', n=24, obs=4, graphs=3, area=2.0, seed=44, init_seed=2)
    sb, env, algo, data = _prepare(meta)
    cbf, act = sd_clone(algo.cbf), sd_clone(algo.actor)
    ob = oracle_batch(sb)
    want = O.update_step('DubinsCar', cbf, act, {}, {}, sb.states, sb.goals, ob['edge_index'], ob['u_ref'], 3, 24,
                         sb.num_obs, K=ob['K'], apply_optim=False)
    algo.train_step(data, apply_optim=False)
    for mod, ref in ((algo.cbf, want['raw_grads']['cbf']), (algo.actor, want['raw_grads']['actor'])):
        total_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in ref.values()))
        err = torch.sqrt(sum(((p.grad.cpu().double() - ref[name].double()) ** 2).sum() for name, p in mod.named_parameters()))
        # one ReLU on/off decision that differs by rounding (pre-activation ~ 0) already costs ~1/sqrt(#units) = 0.3 %
        # of a layer's gradient norm; a wrong kernel costs O(1).  Exactness given identical masks is tested in
        # test_kernels_gpu.py::test_net_backward_exact_given_same_relu_masks.
        assert err / total_ref < 2e-2, (err.item(), total_ref.item())


@pytest.mark.parametrize('case', ['dubins_n16_o4_b3', 'simplecar_c1', 'drone_n8_b2'])
def test_apply_controller_matches_reference(case):
    """GCBF.apply (test-time controller) with the noise switched off against the reference's own apply():
    up to 31 Adam(lr=0.1) iterations through forward_graph -> CBF; Adam's normalised step amplifies rounding in
    near-zero gradient components, hence the looser tolerance."""
    fix = load_golden(case)
    sb, env, algo, data = _prepare(fix['meta'], case)
    n, N = sb.num_agents, sb.nodes_per_graph
    single = env.graph_from_states(sb.states[:N].to(DEV))
    a = algo.apply(single, rand=0)
    assert a.shape == fix['apply_action'].shape
    err = (a.cpu() - fix['apply_action']).abs().max().item()
    assert err <= 2e-3 * max(1.0, fix['apply_action'].abs().max().item()), err


def test_module_api_matches_reference_signatures():
    """CBFGNNLayer.forward(x, edge_attr, edge_index) -> [N, output_dim] on ALL nodes; attention(data) -> [E, 1]."""
    meta = dict(env='DubinsCar', n=16, obs=4, graphs=2, area=2.0, seed=45)
    sb, env, algo, data = _prepare(meta)
    layer = algo.cbf.feat_transformer.module_0
    with torch.no_grad():
        cbf_sd = sd_clone(algo.cbf)
        out = layer(data.x, data.edge_attr, data.edge_index)
        ob = oracle_batch(sb)
        want = O.gnn_layer(cbf_sd, 'feat_transformer.module_0', ob['x'], O.edge_attr('DubinsCar', sb.states, ob['edge_index']),
                           ob['edge_index'], True)
    assert out.shape == (data.x.shape[0], 1024)
    assert (out.cpu() - want).abs().max().item() <= 2e-5
    att = algo.cbf.attention(data)
    assert att.shape == (data.edge_index.shape[1], 1)
    sums = torch.zeros(data.x.shape[0], device=DEV).index_add(0, data.edge_index[1], att.reshape(-1))
    has = torch.bincount(data.edge_index[1], minlength=data.x.shape[0]) > 0
    assert torch.allclose(sums[has], torch.ones_like(sums[has]), atol=1e-5)


def test_update_api_with_buffer_and_batch_collation():
    """GCBF.step / update through the reference-shaped loop: graphs appended one by one, collated by Batch."""
    meta = dict(env='SimpleCar', n=8, obs=0, graphs=1, area=1.5, seed=46)
    sb, env, algo, data = _prepare(meta)
    algo.batch_size = 20
    algo.params['inner_iter'] = 2
    for k in range(8):
        sbk = synth.make_states('SimpleCar', 8, 0, 1, 1.5, 100 + k)
        g = env.graph_from_states(sbk.states.to(DEV))
        a = algo.step(g, prob=0.0)
        assert a.shape == (8, 2)

    class W:
        def __init__(self):
            self.tags = []

        def add_scalar(self, tag, val, it):
            self.tags.append(tag)
            assert val == val
    w = W()
    info = algo.update(1, w)
    assert set(info) == {'acc/safe', 'acc/unsafe', 'acc/derivative'} and len(w.tags) == 14
    assert algo.buffer.size == 0 and algo.memory.size == 8


def test_full_size_properties_c2():
    """BASELINE config C2 (SimpleCar n=256, B=32): properties that do not need the oracle at this size."""
    c = synth.CONFIGS['C2']
    meta = dict(env=c['env'], n=c['num_agents'], obs=c['num_obs'], graphs=c['num_graphs'], area=c['area_size'], seed=c['seed'])
    sb, env, algo, data = _prepare(meta)
    ei = data.edge_index
    key = ei[1] * (ei.max() + 1) + ei[0]
    assert (key[1:] > key[:-1]).all()                                   # strictly sorted (target, source): no duplicates
    rev = torch.stack([ei[1], ei[0]])
    keyr = torch.sort(rev[1] * (ei.max() + 1) + rev[0])[0]
    assert torch.equal(keyr, key)                                       # SimpleCar graphs are symmetric
    d = (sb.states.to(DEV)[ei[0], :2] - sb.states.to(DEV)[ei[1], :2]).norm(dim=1)
    assert d.max() < 1.0 + 1e-6 and (ei[0] // 256 == ei[1] // 256).all()  # within radius, within graph
    res = algo.train_step(data)
    s = res['scalars']
    assert torch.isfinite(s).all() and s[7].item() == 256 * 32
    # linearity of the loss kernels: scalars[6] == sum coef * loss
    hp = algo.params
    tot = hp['loss_unsafe_coef'] * s[0] + hp['loss_safe_coef'] * s[1] + hp['loss_h_dot_coef'] * s[2] + hp['loss_action_coef'] * s[3]
    assert abs(tot.item() - s[6].item()) < 1e-6
    # exact pair count agrees with a brute-force M x M on this size (8192^2 booleans = 64 MB)
    hdot, h = res['hdot'], res['h'].reshape(-1)
    brute = ((hdot.unsqueeze(0) + hp['alpha'] * h.unsqueeze(1)) >= 0).float().mean()
    assert abs(brute.item() - float(res['acc_h_dot'])) < 1e-6


def test_side_stream_overlap_changes_nothing(monkeypatch):
    """GCBF.train_step with the actor / re-linked passes on a side stream (default for small batches) against the same
    step on a single stream: same edges, losses, outputs, power-iteration state and accumulated gradients.  Kernels and
    operand order are identical; only split-K / colsum atomics may reorder, hence 1e-6 / 1e-5 instead of bit equality.
    (One step, before the optimizer: Adam's first update is +-lr for every entry whatever its magnitude, so a rounding-level
    difference in a near-zero gradient entry flips a weight by 2 lr and a second step would no longer be comparable.)"""
    meta = dict(env='DubinsCar', n=64, obs=8, graphs=6, area=4.0, seed=77)
    outs = []
    for mode in ('0', '1'):
        monkeypatch.setenv('GCBF_TWO_STREAMS', mode)
        sb, env, algo, data = _prepare(meta)
        res = algo.train_step(data, apply_optim=False)
        torch.cuda.synchronize()
        u = algo.cbf.state_dict()
        outs.append(dict(s=res['scalars'].clone(), h=res['h'].clone(), hn=res['h_next_new'].clone(), ei=res['edge_index_new'].clone(),
                         g=algo._bucket.grad.clone(), uv=[v.clone() for k, v in u.items() if k.endswith(('_u', '_v'))]))
    a, b = outs
    assert torch.equal(a['ei'], b['ei'])
    assert torch.allclose(a['s'], b['s'], rtol=0, atol=1e-6)
    assert torch.allclose(a['h'], b['h'], rtol=0, atol=1e-6) and torch.allclose(a['hn'], b['hn'], rtol=0, atol=1e-6)
    assert len(a['uv']) > 0
    for x, y in zip(a['uv'], b['uv']):
        assert torch.equal(x, y)                       # three power iterations per step, in program order, on either layout
    assert a['g'].norm() > 0 and (a['g'] - b['g']).norm() <= 1e-5 * a['g'].norm()


def test_full_size_properties_c3_graph_and_masks():
    """BASELINE config C3 (DubinsCar n=1024, 32 obstacles, B=64) at full size: graph and mask invariants that need no
    oracle -- targets are agents only, sorted (target, source) without duplicates or self loops, within radius and within
    graph; an agent within 2R of anything is unsafe and never safe; masks are disjoint."""
    c = synth.CONFIGS['C3']
    meta = dict(env=c['env'], n=c['num_agents'], obs=c['num_obs'], graphs=8, area=c['area_size'], seed=c['seed'])
    sb, env, algo, data = _prepare(meta)
    ei = data.edge_index
    N = sb.nodes_per_graph
    key = ei[1] * (ei.max() + 1) + ei[0]
    assert (key[1:] > key[:-1]).all() and (ei[0] != ei[1]).all()
    assert ((ei[1] % N) < sb.num_agents).all() and (ei[0] // N == ei[1] // N).all()
    st = sb.states.to(DEV)
    d = (st[ei[0], :2] - st[ei[1], :2]).norm(dim=1)
    assert d.max() < env._params['comm_radius'] + 1e-6
    safe, unsafe = env.safe_mask(data), env.unsafe_mask(data)
    assert not (safe & unsafe).any()
    R = env._params['car_radius']
    pos = st[:, :2].reshape(8, N, 2)
    dist = torch.cdist(pos[:, :sb.num_agents], pos) + torch.eye(N, device=DEV)[:sb.num_agents].unsqueeze(0) * 1e6
    close = (dist.min(dim=2).values < 2 * R).reshape(-1)
    assert (unsafe[close]).all() and not (safe[close]).any()


@pytest.mark.parametrize('env_name,n,obs,area', [('SimpleCar', 12, 0, 2.0), ('DubinsCar', 10, 4, 2.0), ('SimpleDrone', 6, 6, 1.0)])
def test_device_replay_collates_like_from_data_list(env_name, n, obs, area):
    """SURVEY 8f-2 (replay half): a batch gathered from the device-resident ring and re-linked by the batched graph kernels
    is the batch `Batch.from_data_list` builds from the stored `Data` objects -- same edges (bit-exact), edge features,
    node types, states and nominal controls -- and GCBF.update runs on it."""
    import random
    import numpy as np
    from gcbf_b200.algo.device_buffer import collate
    from gcbf_b200.data import Batch
    meta = dict(env=env_name, n=n, obs=obs, graphs=1, area=area, seed=61)
    sb, env, algo, data = _prepare(meta)
    _, _, algo_ring, _ = _prepare(meta)
    algo_ring._env = env
    algo_ring.use_device_replay(capacity=4)
    for k in range(14):
        sbk = synth.make_states(env_name, n, obs, 1, area, 300 + k)
        g = env.graph_from_states(sbk.states.to(DEV))
        algo.buffer.append(g, is_safe=(k % 3 != 0))
        algo_ring.buffer.append(g, is_safe=(k % 3 != 0))
    for seed, (cnt, m, bal) in enumerate([(5, 3, False), (6, 3, True)]):
        np.random.seed(seed), random.seed(seed)
        want = Batch.from_data_list(algo.buffer.sample(cnt, m, bal))
        np.random.seed(seed), random.seed(seed)
        got = collate(env, [(algo_ring.buffer, algo_ring.buffer.sample(cnt, m, bal))])
        assert torch.equal(got.edge_index, want.edge_index)
        for key in ('states', 'u_ref', 'edge_attr', 'x'):
            assert torch.equal(getattr(got, key), getattr(want, key)), key
        if hasattr(want, 'agent_mask'):
            assert torch.equal(got.agent_mask, want.agent_mask)
    algo_ring.batch_size = 20
    algo_ring.params['inner_iter'] = 2
    np.random.seed(3), random.seed(3)
    info = algo_ring.update(1, None)
    assert set(info) == {'acc/safe', 'acc/unsafe', 'acc/derivative'} and all(v == v for v in info.values())
    assert algo_ring.buffer.size == 0 and algo_ring.memory.size == 14
    np.random.seed(4), random.seed(4)
    for k in range(6):
        sbk = synth.make_states(env_name, n, obs, 1, area, 400 + k)
        algo_ring.buffer.append(env.graph_from_states(sbk.states.to(DEV)), is_safe=(k % 2 == 0))
    info = algo_ring.update(2, None)                           # balanced sampling from the fresh buffer and the merged memory
    assert all(v == v for v in info.values()) and algo_ring.memory.size == 20
