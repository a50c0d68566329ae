"""GPU parity tests of the MACBF path through the C ABI: the CUDA kernels of csrc/macbf.cu and the product's MACBF train
step against the reference-on-shim fixtures (tests/golden/macbf_*.pt) and the CPU port (oracle/macbf_oracle.py).
Tolerances: edge lists and masks bit-exact; h, u, losses <= 1e-5 absolute; post-Adam weights as in the GCBF tests."""
import copy

import pytest
import torch

import gcbf_oracle as O
import macbf_oracle as MO
from conftest import digest_close, load_golden, macbf_golden_cases
from helpers import case_inputs

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')
CASES = macbf_golden_cases()


def _inputs(fix):
    meta = fix['meta']
    sb = case_inputs(meta)
    if meta['case'].endswith('single'):
        sb.states[0, :2] = sb.goals[0, :2]
    assert torch.equal(sb.states, fix['states'])
    return meta, sb


def _env_algo(sb, fix=None, k=12):
    from gcbf_b200.algo import MACBF
    from gcbf_b200.env import make_env
    env = make_env(sb.env, sb.num_agents, DEV)
    params = env.default_params
    params.update({'num_obs': sb.num_obs, 'area_size': sb.area_size})
    env = make_env(sb.env, sb.num_agents, DEV, params=params, max_neighbors=k)
    algo = MACBF(env, sb.num_agents, env.node_dim, env.edge_dim, env.action_dim, DEV, 512, MO.HYPERPARAMS[sb.env], reference_rng=False)
    if fix is not None:
        algo.cbf.load_state_dict(fix['cbf_init'])
        algo.actor.load_state_dict(fix['actor_init'])
    return env, algo


@pytest.mark.parametrize('into_param', [True, False])
@pytest.mark.parametrize('case', CASES)
def test_macbf_forward_and_train_steps_match_the_reference(case, into_param):
    from gcbf_b200 import synth
    fix = load_golden(case)
    meta, sb = _inputs(fix)
    env, algo = _env_algo(sb, fix)
    algo.GRAD_INTO_PARAM = into_param        # weight gradients accumulated by the kernels / returned through autograd
    data = synth.product_batch(env, sb, DEV)
    assert torch.equal(data.edge_index.cpu(), fix['edge_index'])                        # top-k filtered radius graph: bit-exact
    assert torch.allclose(data.edge_attr.cpu(), fix['edge_attr'], rtol=0, atol=1e-6)
    with torch.no_grad():
        h, u = algo.cbf(data), algo.act(data)
    assert (h.cpu() - fix['h_probe']).abs().max().item() <= 1e-5
    assert (u.cpu() - fix['u_probe']).abs().max().item() <= 1e-5
    assert torch.equal(env.safe_mask(data, return_edge=True).cpu(), fix['safe_mask'])
    assert torch.equal(env.unsafe_mask(data, return_edge=True).cpu(), fix['unsafe_mask'])
    counts = {'acc/unsafe': int(fix['unsafe_mask'].sum()), 'acc/safe': int(fix['safe_mask'].sum()), 'acc/derivative': int(fix['edge_index'].shape[1])}
    for gold in fix['steps']:
        s = algo.train_step(data)['scalars'].tolist()
        got = {'loss/unsafe': s[0], 'loss/safe': s[1], 'loss/derivative': s[2], 'loss/action': s[3], 'acc/unsafe': s[4], 'acc/safe': s[5],
               'acc/derivative': s[7]}
        for tag, val in got.items():
            tol = 1e-5 if tag.startswith('loss') else 1.5 / max(1, counts[tag])          # accuracies: one flipped sample
            assert abs(val - gold['scalars'][tag]) <= tol, (tag, val, gold['scalars'][tag])
    # post-step weights after two clipped Adam steps: digests within 1e-5 plus the GCBF tests' budget of 1 % of the elements taking a
    # sign-flipped first Adam step (conftest.digest_close)
    sd_c = {k: v.detach().cpu() for k, v in algo.cbf.state_dict().items()}
    sd_a = {k: v.detach().cpu() for k, v in algo.actor.state_dict().items()}
    bad = digest_close(sd_c, fix['cbf_final'], 1e-5, 1e-5, flip=0.01 * 2 * 3e-4 * 2)
    assert not bad, bad[:3]
    bad = digest_close(sd_a, fix['actor_final'], 1e-5, 1e-5, flip=0.01 * 2 * 1e-3 * 2)
    assert not bad, bad[:3]


def test_macbf_gradients_match_the_port():
    """Raw gradients of one step (no optimiser) against autograd on the CPU port, per net relative to the net's gradient norm."""
    from gcbf_b200 import synth
    fix = load_golden('macbf_dubins_n24_o6_b3')
    meta, sb = _inputs(fix)
    env, algo = _env_algo(sb, fix)
    data = synth.product_batch(env, sb, DEV)
    res = algo.train_step(data, apply_optim=False)
    cbf, act = copy.deepcopy(fix['cbf_init']), copy.deepcopy(fix['actor_init'])
    want = MO.update_step(sb.env, cbf, act, {}, {}, sb.states, sb.goals, fix['edge_index'], fix['u_ref'], sb.num_graphs, sb.num_agents,
                          sb.num_obs, apply_optim=False)
    assert (res['h'].cpu() - want['h']).abs().max().item() <= 1e-5 and (res['h_next'].cpu() - want['h_next']).abs().max().item() <= 1e-5
    for mod, key in ((algo.cbf, 'cbf'), (algo.actor, 'actor')):
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in want['raw_grads'][key].values()))
        for name, p in mod.named_parameters():
            diff = (p.grad.cpu().double() - want['raw_grads'][key][name].double()).norm()
            assert diff <= 1e-4 * total + 1e-12, (key, name, float(diff), float(total))


@pytest.mark.parametrize('env_name,n,o,B,area,k', [('DubinsCar', 256, 32, 8, 3.0, 12), ('SimpleDrone', 64, 64, 3, 0.9, 12), ('SimpleCar', 300, 0, 4, 2.5, 12),
                                                    ('DubinsCar', 16, 0, 2, 1.0, 3), ('SimpleCar', 9, 0, 2, 0.5, 2), ('SimpleDrone', 14, 14, 1, 5.0, 12)])
def test_radius_graph_topk_bit_exact(env_name, n, o, B, area, k):
    from gcbf_b200 import synth
    sb = synth.make_states(env_name, n, o, B, area, 77)
    env, _ = _env_algo(sb, None, k)
    data = synth.product_batch(env, sb, DEV)
    want = MO.batch_radius_graph_topk(env_name, sb.states, B, sb.nodes_per_graph, n, k)
    assert torch.equal(data.edge_index.cpu(), want)
    sm, um = MO.edge_masks(env_name, O.edge_attr(env_name, sb.states, want))
    got = env.edge_masks(data).cpu()
    # edge features are computed on the device (cos / sin for DubinsCar): compare the masks where the distance is not within 1e-6 of a threshold
    dist = O.edge_attr(env_name, sb.states, want)[:, :O.ENV_PARAMS[env_name]['pos_dim']].norm(dim=-1)
    R = O.ENV_PARAMS[env_name]['radius']
    clear = ((dist - 4 * R).abs() > 1e-6) & ((dist - 2 * R).abs() > 1e-6)
    assert torch.equal(got[0][clear], sm[clear]) and torch.equal(got[1][clear], um[clear])


@pytest.mark.parametrize('C,deg_hi,ld_pad', [(128, 13, 0), (5, 40, 3), (1, 3, 0)])
def test_seg_max_fwd_bwd(C, deg_hi, ld_pad):
    from gcbf_b200 import ops
    g = torch.Generator().manual_seed(C)
    Nn = 57
    deg = torch.randint(0, deg_hi + 1, (Nn,), generator=g)
    deg[3] = 0
    deg[Nn - 1] = 0
    dst = torch.repeat_interleave(torch.arange(Nn), deg)
    E = int(deg.sum())
    rowptr = torch.zeros(Nn + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(deg, 0).int()
    buf = torch.randn(E, C + ld_pad, generator=g)
    msg_cpu = buf[:, :C].clone().requires_grad_(True)
    want = torch.zeros(Nn, C).scatter_reduce(0, dst.view(-1, 1).expand(E, C), msg_cpu, reduce='amax', include_self=False)
    d_out = torch.randn(Nn, C, generator=g)
    want.backward(d_out)
    msg = buf.to(DEV)[:, :C].requires_grad_(True)                     # strided rows when ld_pad > 0
    out = ops.SegMaxFunction.apply(msg, rowptr.to(DEV), Nn)
    out.backward(d_out.to(DEV))
    assert torch.equal(out.detach().cpu(), want.detach())
    assert torch.equal(msg.grad.cpu(), msg_cpu.grad)


def test_seg_max_on_a_graph_without_edges():
    from gcbf_b200 import ops
    rowptr = torch.zeros(6, dtype=torch.int32, device=DEV)
    msg = torch.empty(0, 7, device=DEV, requires_grad=True)
    out = ops.SegMaxFunction.apply(msg, rowptr, 5)
    out.sum().backward()
    assert out.shape == (5, 7) and float(out.detach().abs().max()) == 0.0 and msg.grad.shape == (0, 7)


@pytest.mark.parametrize('E,M,ad,empty', [(5000, 600, 2, None), (64, 10, 3, 'unsafe'), (33, 7, 2, 'safe'), (1, 1, 2, None)])
def test_macbf_loss_kernels_match_autograd(E, M, ad, empty):
    from gcbf_b200 import _C
    g = torch.Generator().manual_seed(E)
    h = (torch.randn(E, generator=g) * 0.05).requires_grad_(True)
    hn = (h.detach() + torch.randn(E, generator=g) * 0.002).requires_grad_(True)
    act = torch.randn(M, ad, generator=g).requires_grad_(True)
    safe = torch.rand(E, generator=g) < 0.6
    unsafe = (torch.rand(E, generator=g) < 0.2) & ~safe
    if empty == 'unsafe':
        unsafe[:] = False
    if empty == 'safe':
        safe[:] = False
    alpha, eps, dt, cu, cs, ch, ca = 1.0, 0.02, 0.03, 1.0, 0.7, 0.4, 0.05
    hu, hs = h[unsafe], h[safe]
    lu = torch.relu(hu + eps).mean() if hu.numel() else torch.tensor(0.0)
    ls = torch.relu(-hs + eps).mean() if hs.numel() else torch.tensor(0.0)
    h_dot = (hn - h) / dt
    lh = torch.relu(-h_dot - alpha * h + eps).mean()
    la = torch.square(act).sum(dim=1).mean()
    loss = cu * lu + cs * ls + ch * lh + ca * la
    loss.backward()
    acc = [float((hu < 0).float().mean()) if hu.numel() else 1.0, float((hs >= 0).float().mean()) if hs.numel() else 1.0,
           float(((h_dot + alpha * h) >= 0).float().mean())]
    hd, hnd, ad_ = h.detach().to(DEV), hn.detach().to(DEV), act.detach().to(DEV).contiguous()
    s8, u8 = safe.to(torch.uint8).to(DEV), unsafe.to(torch.uint8).to(DEV)
    partial = torch.empty(16, device=DEV, dtype=torch.float64)
    d_h, d_hn, d_act, sc = torch.empty(E, device=DEV), torch.empty(E, device=DEV), torch.empty(M, ad, device=DEV), torch.empty(8, device=DEV)
    _C.call('gcbf_macbf_loss_partials', _C.ptr(hd), _C.ptr(hnd), _C.ptr(s8), _C.ptr(u8), E, _C.ptr(ad_), ad, M, alpha, eps, dt, _C.ptr(partial))
    _C.call('gcbf_macbf_loss_grads', _C.ptr(hd), _C.ptr(hnd), _C.ptr(s8), _C.ptr(u8), E, _C.ptr(ad_), ad, M, alpha, eps, dt, cu, cs, ch, ca,
            _C.ptr(partial), _C.ptr(d_h), _C.ptr(d_hn), _C.ptr(d_act), _C.ptr(sc))
    for got, want in zip(sc.tolist(), [float(lu), float(ls), float(lh), float(la), acc[0], acc[1], float(loss), acc[2]]):
        assert abs(got - want) <= 1e-6
    assert torch.allclose(d_h.cpu(), h.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(d_hn.cpu(), hn.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(d_act.cpu(), act.grad, rtol=1e-6, atol=1e-9)
    assert partial[7].item() == E and partial[10].item() == M


def test_macbf_rollout_step_and_update_api():
    """MACBF.step fills the buffer from env steps, update() runs `inner_iter` train steps on sampled segments, save / load round-trip."""
    import random
    import tempfile
    import numpy as np
    from gcbf_b200 import synth
    random.seed(0)
    np.random.seed(0)                       # the buffer samples its windows with the host RNGs, like the reference
    sb = synth.make_states('DubinsCar', 16, 4, 1, 2.5, 5)
    env, algo = _env_algo(sb)
    algo.params['inner_iter'] = 2
    algo.batch_size = 20
    env.set_goal(sb.goals)
    env._obs = sb.obs.to(DEV)
    data = env.graph_from_states(sb.states.to(DEV))
    env._data, env._t = data, 0
    for _ in range(8):
        data.update(type(data)(u_ref=env.u_ref(data)))
        a = algo.step(data, prob=0.0)
        assert a.shape == (16, 2)
        data, _, _, _ = env.step(a)
    data.update(type(data)(u_ref=env.u_ref(data)))
    assert algo.buffer.size == 8
    info = algo.update(0)
    assert set(info) == {'acc/safe', 'acc/unsafe', 'acc/derivative'} and all(0.0 <= v <= 1.0 for v in info.values())
    assert algo.buffer.size == 0 and algo.memory.size == 8
    with tempfile.TemporaryDirectory() as d:
        algo.save(d)
        before = algo.act(data).clone()
        for p in algo.actor.parameters():
            p.data.add_(0.01)
        algo.load(d)
        assert torch.equal(algo.act(data), before)
    assert torch.equal(algo.apply(data), algo.act(data))
