"""MACBF baseline on the GPU: train-step time (forward + losses + backward + clip + Adam), graph build with the top-12
filter and the actor's rollout-time latency on synthetic batches, with the CPU port (oracle/macbf_oracle.py) timed beside it on a
bounded sample.  Device-timed with CUDA events after warm-up.
    python tools/macbf_probe.py [out.json]           # both workloads
    python tools/macbf_probe.py --one-step C3        # one train step only (for an ncu capture of the same command)
"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), os.path.join(ROOT, 'oracle'), ROOT]
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
        sb, env, algo = build(name, dev)
        st = sb.states.to(dev)
        data = synth.product_batch(env, sb, dev)
        E, M = int(data.edge_index.shape[1]), int(data.u_ref.shape[0])
        _C.reset_counters()
        algo.train_step(data)
        launches = _C.kernel_launches()
        ms_step = timed(lambda: algo.train_step(data), 20, 5)
        ms_graph = timed(lambda: env.graph_from_states(st), 20, 3)
        with torch.no_grad():
            ms_act = timed(lambda: algo.act(data), 20, 3)
        rec = dict(workload=f"{sb.env} n={sb.num_agents} obs={sb.num_obs} B={sb.num_graphs} area={sb.area_size} max_neighbors=12", agents=M, edges=E,
                   max_in_degree=int(torch.bincount(data.edge_index[1]).max()), train_step_ms=round(ms_step, 4),
                   agent_steps_per_s=round(M / ms_step * 1e3, 1), graph_build_ms=round(ms_graph, 4), actor_forward_ms=round(ms_act, 4),
                   gpu_launches_per_step=int(launches), loss=float(algo.train_step(data)['scalars'][6]))
        if name == 'ref':
            # CPU port on the same batch (a bounded sample: 3 steps), all host threads the container may use
            import gcbf_oracle as O, macbf_oracle as MO
            import bench
            torch.set_num_threads(bench.host_threads())
            cbf = {k: v.detach().cpu().clone() for k, v in algo.cbf.state_dict().items()}
            act = {k: v.detach().cpu().clone() for k, v in algo.actor.state_dict().items()}
            ei, ur = data.edge_index.cpu(), data.u_ref.cpu()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                MO.update_step(sb.env, cbf, act, {}, {}, sb.states, sb.goals, ei, ur, sb.num_graphs, sb.num_agents, sb.num_obs)
                ts.append(time.perf_counter() - t0)
            rec['cpu_port'] = dict(seconds_per_step=round(sorted(ts)[1], 4), agent_steps_per_s=round(M / sorted(ts)[1], 1), cores=torch.get_num_threads(),
                                   sample='the same batch, median of 3 steps, graph given')
        out[name] = rec
        print(json.dumps({name: rec}))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], 'w'), indent=1)


if __name__ == '__main__':
    main()
