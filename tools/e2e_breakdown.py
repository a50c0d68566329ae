import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gcbf-pytorch_b200')); sys.path.insert(0, ROOT)
from gcbf_b200 import ops, _C
from gcbf_b200.data import Data
import bench
dev = torch.device('cuda:0')
sb, env, algo = bench.build_case('C2', dev, 0)
data = env.graph_from_states(sb.states.to(dev))
for _ in range(3): algo.train_step(data)
torch.cuda.synchronize()
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f'{label:28s} {(time.perf_counter()-t0)*1e3:8.2f} ms', flush=True); return r
hs = sb.states.pin_memory()
for it in range(2):
    st = T('h2d', lambda: hs.to(dev, non_blocking=True))
    g = T('make_graph', lambda: env.make_graph(st))
    ei = T('radius_graph', lambda: ops.radius_graph(g.states, 2, 32, 256, 256, 1.0, 0)[0])
    ea = T('edge_attr', lambda: env.edge_attr(g.states, ei))
    g.update(Data(edge_index=ei, edge_attr=ea))
    ur = T('u_ref', lambda: env.u_ref(g))
    g.update(Data(u_ref=ur))
    T('train_step(new graph)', lambda: algo.train_step(g))
    T('train_step(new graph) again', lambda: algo.train_step(g))
    T('train_step(old graph)', lambda: algo.train_step(data))
    T('graph_from_states', lambda: env.graph_from_states(st))
