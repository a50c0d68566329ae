"""MACBF baseline on the GPU: train-step time (forward + losses + backward + clip + Adam), graph build with the top-12
filter and the actor's rollout-time latency on synthetic batches.  Device-timed with CUDA events after warm-up.  (`bench.py --macbf`
prints the same measurement as a bench line with the CPU port timed beside it; this tool never touches oracle/.)
    python tools/macbf_probe.py [out.json]           # both workloads
    python tools/macbf_probe.py --one-step C3        # two train steps only (for an ncu capture of the same command)
"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), ROOT]
from gcbf_b200 import _C, synth
from gcbf_b200.algo import MACBF
from gcbf_b200.env import make_env
from gcbf_b200.trainer.utils import read_params

# the reference's own training scale (16 agents per graph, a batch of ~300 sampled graphs, macbf.py:126-133) and the C3 shape
WORKLOADS = {
    'ref': dict(env='DubinsCar', num_agents=16, num_obs=8, num_graphs=256, area_size=4.0, seed=2001),
    'C3': dict(env='DubinsCar', num_agents=1024, num_obs=32, num_graphs=64, area_size=32.0, seed=1003),
}


def build(name, dev):
    c = WORKLOADS[name]
    sb = synth.make_states(c['env'], c['num_agents'], c['num_obs'], c['num_graphs'], c['area_size'], c['seed'])
    env = make_env(sb.env, sb.num_agents, dev)
    params = env.default_params
    params.update({'num_obs': sb.num_obs, 'area_size': sb.area_size})
    env = make_env(sb.env, sb.num_agents, dev, params=params, max_neighbors=12)
    torch.manual_seed(0)
    algo = MACBF(env, sb.num_agents, env.node_dim, env.edge_dim, env.action_dim, dev, 512, read_params(sb.env, 'macbf'), reference_rng=False)
    return sb, env, algo


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def measure(name, dev, steps=20, warmup=5):
    """(record, (sb, algo, data)) for one workload."""
    sb, env, algo = build(name, dev)
    st = sb.states.to(dev)
    data = synth.product_batch(env, sb, dev)
    E, M = int(data.edge_index.shape[1]), int(data.u_ref.shape[0])
    _C.reset_counters()
    algo.train_step(data)
    launches = _C.kernel_launches()
    ms_step = timed(lambda: algo.train_step(data), steps, warmup)
    ms_graph = timed(lambda: env.graph_from_states(st), steps, 3)
    with torch.no_grad():
        ms_act = timed(lambda: algo.act(data), steps, 3)
    rec = dict(workload=f"{sb.env} n={sb.num_agents} obs={sb.num_obs} B={sb.num_graphs} area={sb.area_size} max_neighbors=12", agents=M, edges=E,
               max_in_degree=int(torch.bincount(data.edge_index[1]).max()), train_step_ms=round(ms_step, 4),
               agent_steps_per_s=round(M / ms_step * 1e3, 1), graph_build_ms=round(ms_graph, 4), actor_forward_ms=round(ms_act, 4),
               gpu_launches_per_step=int(launches), loss=float(algo.train_step(data)['scalars'][6]))
    return rec, (sb, algo, data)


def main():
    dev = torch.device('cuda', 0)
    if len(sys.argv) > 1 and sys.argv[1] == '--one-step':
        sb, env, algo = build(sys.argv[2] if len(sys.argv) > 2 else 'C3', dev)
        data = synth.product_batch(env, sb, dev)
        algo.train_step(data)
        torch.cuda.synchronize()
        algo.train_step(data)
        torch.cuda.synchronize()
        return
    out = {}
    for name in WORKLOADS:
        out[name] = measure(name, dev)[0]
        print(json.dumps({name: out[name]}))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
